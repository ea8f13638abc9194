#!/usr/bin/env python
"""A few device-resident steps of the hot path on one GPU, small enough to sit under ncu:
    python tools/prof_step.py [--reads N] [--reps R] [--workload rps12|long|noisy] [--flags dp|full]
Prints the CUDA-event time per step (never quote a number printed under a profiler)."""
import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=400_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--workload", default="rps12", choices=["rps12", "long", "noisy"])
    ap.add_argument("--flags", default="", help="dp: DP on every read; full: every ops row")
    a = ap.parse_args()
    import torch
    import treat_b200 as T
    from treat_b200 import cabi, synth
    noise = {}
    if a.workload == "long":
        t, p_trunc, seed = synth.long_template(300, 8), 0.2, synth.SEED_BASE + 5
    else:
        t, p_trunc, seed = synth.rps12_template(), 0.0, synth.SEED_BASE + 3
        if a.workload == "noisy":
            noise = {"p_sub_base": 0.01, "p_indel_read": 0.05}
    with tempfile.TemporaryDirectory() as d:
        tm = T.NewTemplateFromFasta(t.write_fasta(os.path.join(d, "t.fa")), T.FORWARD, t.edit_base)
    eng = T.Aligner(0)
    eng.SetTemplate(tm)
    N = a.reads
    seqs, off, rc = synth.reads(t, N, seed=seed, p_trunc=p_trunc, **noise)
    max_len = int(np.diff(off.astype(np.int64)).max())
    stride = eng.ops_stride(max_len)
    dev = torch.device("cuda", 0)
    d_seqs = torch.from_numpy(np.ascontiguousarray(seqs)).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    d_rc = torch.from_numpy(rc.astype(np.int32)).to(dev)
    d_rec = torch.empty(N * 48, dtype=torch.uint8, device=dev)
    d_ops = torch.empty(N * stride, dtype=torch.uint8, device=dev)
    flags = cabi.WANT_OPS | cabi.OPS_GAPPED_ONLY
    if "dp" in a.flags:
        flags |= cabi.DP_EVERY_READ
    if "full" in a.flags:
        flags &= ~cabi.OPS_GAPPED_ONLY
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)

    def step():
        eng.align_batch_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rc.data_ptr(), N, max_len, flags, d_rec.data_ptr(),
                               d_ops.data_ptr(), stride, st.cuda_stream)
    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(a.reps):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    print("reads that needed a DP per step: pure deletions %d, banded %d, full matrix %d" % tuple(v // (a.reps + 1) for v in eng.dp_stats()))
    print("%s: %d reads, %.3f ms per step (%.1f M reads/s), %d launches" % (
        a.workload, N, e0.elapsed_time(e1) / a.reps, N / (e0.elapsed_time(e1) / a.reps) / 1e3, eng.launch_count()))
    eng.close()


if __name__ == "__main__":
    main()
