#!/usr/bin/env python
"""What the host's PCIe / memory path gives N GPUs at once: plain pinned cudaMemcpyAsync H2D copies of a step's input bytes
(treat_b200_measure_h2d) on GPUs 0..N-1 concurrently, from N threads of one process (--mode threads, the device-group form)
or from N processes (--mode procs, the one-process-per-GPU form).  One JSON line per configuration."""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(dev, nbytes, reps, barrier, out, idx):
    from treat_b200 import cabi
    L = cabi.lib()
    ctx = C.c_void_p()
    cabi.check(L.treat_b200_create(dev, C.byref(ctx)))
    hp = C.c_void_p()
    cabi.check(L.treat_b200_host_alloc(C.byref(hp), nbytes))
    C.memset(hp, 65, nbytes)
    ms = C.c_double()
    cabi.check(L.treat_b200_measure_h2d(ctx, hp, nbytes, 2, C.byref(ms)), ctx)  # warm-up
    barrier.wait()
    t0 = time.perf_counter()
    cabi.check(L.treat_b200_measure_h2d(ctx, hp, nbytes, reps, C.byref(ms)), ctx)
    out[idx] = ms.value
    barrier.wait()
    L.treat_b200_host_free(hp)
    L.treat_b200_destroy(ctx)
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--mbytes", type=int, default=305)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--mode", default="threads,procs")
    a = ap.parse_args()
    nbytes = a.mbytes * 1000 * 1000
    import torch
    have = torch.cuda.device_count()
    for mode in a.mode.split(","):
        for n in [int(x) for x in a.gpus.split(",")]:
            if n > have:
                continue
            if mode == "threads":
                bar = threading.Barrier(n)
                out = [0.0] * n
                th = [threading.Thread(target=one, args=(i, nbytes, a.reps, bar, out, i)) for i in range(n)]
                [t.start() for t in th]
                [t.join() for t in th]
            else:
                ctxm = mp.get_context("spawn")
                bar = ctxm.Barrier(n)
                out = ctxm.Array("d", n)
                ps = [ctxm.Process(target=one, args=(i, nbytes, a.reps, bar, out, i)) for i in range(n)]
                [p.start() for p in ps]
                [p.join() for p in ps]
                out = list(out)
            worst = max(out)
            print(json.dumps({"mode": mode, "gpus": n, "mbytes": a.mbytes, "ms_per_copy_per_gpu": [round(x, 3) for x in out],
                              "worst_ms": round(worst, 3), "gb_per_s_per_gpu_worst": round(nbytes / worst / 1e6, 2),
                              "gb_per_s_all": round(n * nbytes / worst / 1e6, 2)}), flush=True)


if __name__ == "__main__":
    main()
