#!/usr/bin/env python
"""End-to-end call time of treat_b200_align_batch with several GPUs busy at once (diagnostic, GPU box only).
One process per GPU, no torch.distributed: every process prepares the bench workload in pinned memory, waits for the
wall-clock time --t0, then calls (or, with --raw, copies the same bytes in both directions) until t0 + 3 s and reports
the mean over the calls that began inside [t0 + 0.5, t0 + 2.5].  TREAT_B200_CHUNK / TREAT_B200_SLOTS are read from the
environment by the library.
    for i in 0 1 2 3; do python tools/e2e_multi.py --gpu $i --t0 $T0 & done; wait"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--t0", type=float, required=True)
    ap.add_argument("--raw", action="store_true")
    ap.add_argument("--reads", type=int, default=1_000_000)
    a = ap.parse_args()
    import torch
    import bench
    import treat_b200 as T
    from treat_b200 import cabi, synth
    torch.cuda.set_device(a.gpu)
    t = synth.rps12_template()
    with tempfile.TemporaryDirectory() as d:
        tm = T.NewTemplateFromFasta(t.write_fasta(os.path.join(d, "t.fa")), T.FORWARD, t.edit_base)
    eng = T.Aligner(a.gpu)
    eng.SetTemplate(tm)
    hold = []

    def alloc(nbytes):
        arr, p = bench.pinned_array(nbytes)
        hold.append(p)
        return arr
    N = a.reads
    seqs, off, rc = synth.reads(t, N, seed=synth.SEED_BASE + 3 + a.gpu, pinned_alloc=alloc)
    stride = eng.ops_stride(int(np.diff(off.astype(np.int64)).max()))
    h_off, h_rc = alloc((N + 1) * 8).view(np.uint64), alloc(N * 4).view(np.uint32)
    h_off[:] = off
    h_rc[:] = rc
    h_rec = alloc(N * 48).view(cabi.RECORD_DTYPE)
    h_ops = alloc(N * stride).reshape(N, stride)
    flags = cabi.WANT_OPS | cabi.OPS_GAPPED_ONLY
    if a.raw:
        nin = len(seqs) + h_off.nbytes + h_rc.nbytes
        src, dst = alloc(nin), alloc(N * 50)
        tin, tout = torch.from_numpy(src), torch.from_numpy(dst)
        d_in = torch.empty(nin, dtype=torch.uint8, device="cuda")
        d_out = torch.empty(N * 50, dtype=torch.uint8, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def call():
            with torch.cuda.stream(s1):
                d_in.copy_(tin, non_blocking=True)
            with torch.cuda.stream(s2):
                tout.copy_(d_out, non_blocking=True)
            s1.synchronize()
            s2.synchronize()
    else:
        def call():
            eng.align_batch(seqs, h_off, h_rc, flags, rec=h_rec, ops=h_ops)
    call()
    call()
    late = time.time() > a.t0
    while time.time() < a.t0:
        time.sleep(0.001)
    ts = []
    while time.time() < a.t0 + 3.0:
        b = time.time()
        call()
        ts.append((b, time.time() - b))
    mid = [d for (b, d) in ts if a.t0 + 0.5 <= b <= a.t0 + 2.5]
    print("gpu %d %s: %.2f ms per call (%d calls in the window)%s" % (a.gpu, "raw copies" if a.raw else "align_batch", 1e3 * float(np.mean(mid)) if mid else -1.0,
                                                                     len(mid), " LATE START" if late else ""), flush=True)


if __name__ == "__main__":
    main()
