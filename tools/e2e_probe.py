"""Where does the end-to-end time of treat_b200_align_batch go?  (diagnostic, GPU box only)

Times, for the bench workload: the raw pinned H2D / D2H copies alone and together, and the host-buffer call
with and without the ops output, for a few chunk sizes (TREAT_B200_CHUNK is read once per process, so the
chunk sweep re-executes this script)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import treat_b200 as T  # noqa: E402
from treat_b200 import cabi, synth  # noqa: E402


def main():
    import tempfile
    import bench
    N = int(os.environ.get("PROBE_N", "1000000"))
    t = synth.rps12_template()
    with tempfile.TemporaryDirectory() as d:
        tm = T.NewTemplateFromFasta(t.write_fasta(os.path.join(d, "tmpl.fa")), T.FORWARD, t.edit_base)
    eng = T.Aligner(0)
    eng.SetTemplate(tm)
    hold = []

    def alloc(nbytes):
        arr, p = bench.pinned_array(nbytes)
        hold.append(p)
        return arr
    seqs, off, rc = synth.reads(t, N, seed=synth.SEED_BASE + 3, pinned_alloc=alloc)
    max_len = int(np.diff(off.astype(np.int64)).max())
    stride = eng.ops_stride(max_len)
    h_seqs = seqs
    h_off, h_rc = alloc((N + 1) * 8).view(np.uint64), alloc(N * 4).view(np.uint32)
    h_off[:] = off
    h_rc[:] = rc
    h_rec = alloc(N * 48).view(cabi.RECORD_DTYPE)
    h_ops = alloc(N * stride).reshape(N, stride)

    def timeit(fn, reps=5):
        fn(); fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return 1e3 * (time.perf_counter() - t0) / reps

    if os.environ.get("PROBE_RAW", "1") == "1":
        tin = torch.from_numpy(h_seqs)
        d_in = torch.empty(len(seqs), dtype=torch.uint8, device="cuda")
        d_out = torch.empty(N * 48, dtype=torch.uint8, device="cuda")
        tout = torch.from_numpy(h_rec.view(np.uint8).reshape(-1))
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def h2d():
            with torch.cuda.stream(s1):
                d_in.copy_(tin, non_blocking=True)
            s1.synchronize()

        def d2h():
            with torch.cuda.stream(s2):
                tout.copy_(d_out, non_blocking=True)
            s2.synchronize()

        def both():
            with torch.cuda.stream(s1):
                d_in.copy_(tin, non_blocking=True)
            with torch.cuda.stream(s2):
                tout.copy_(d_out, non_blocking=True)
            s1.synchronize(); s2.synchronize()
        a, b, c = timeit(h2d), timeit(d2h), timeit(both)
        # full duplex at equal sizes: D2H of as many bytes as the H2D
        d_big = torch.empty(len(seqs), dtype=torch.uint8, device="cuda")
        h_big_raw = alloc(len(seqs))
        t_big = torch.from_numpy(h_big_raw)

        def duplex():
            with torch.cuda.stream(s1):
                d_in.copy_(tin, non_blocking=True)
            with torch.cuda.stream(s2):
                t_big.copy_(d_big, non_blocking=True)
            s1.synchronize(); s2.synchronize()
        dd = timeit(duplex)
        print("full duplex %.0f MB each way: %.2f ms (%.1f GB/s per direction)" % (len(seqs) / 1e6, dd, len(seqs) / dd / 1e6), flush=True)
        print("raw copies: H2D %.0f MB %.2f ms (%.1f GB/s) | D2H %.0f MB %.2f ms (%.1f GB/s) | both %.2f ms" % (
            len(seqs) / 1e6, a, len(seqs) / a / 1e6, N * 48 / 1e6, b, N * 48 / b / 1e6, c), flush=True)
    if os.environ.get("PROBE_RAW", "1") == "1":
        # the copy pattern of align_batch without any kernel: 65,536-read chunks over 4 streams, per chunk
        # H2D reads + offsets + counts, D2H records
        CH = 65536
        streams = [torch.cuda.Stream() for _ in range(4)]
        t_off, t_rc = torch.from_numpy(h_off.view(np.int64)), torch.from_numpy(h_rc.view(np.int32))
        d_off = torch.empty(N + 1, dtype=torch.int64, device="cuda")
        d_rc = torch.empty(N, dtype=torch.int32, device="cuda")
        offs = off.astype(np.int64)

        def chunked(sync_reuse):
            for ci, c0 in enumerate(range(0, N, CH)):
                c1 = min(N, c0 + CH)
                st = streams[ci % 4]
                if sync_reuse:
                    st.synchronize()
                with torch.cuda.stream(st):
                    d_in[offs[c0]:offs[c1]].copy_(tin[offs[c0]:offs[c1]], non_blocking=True)
                    d_off[c0:c1 + 1].copy_(t_off[c0:c1 + 1], non_blocking=True)
                    d_rc[c0:c1].copy_(t_rc[c0:c1], non_blocking=True)
                    tout[c0 * 48:c1 * 48].copy_(d_out[c0 * 48:c1 * 48], non_blocking=True)
            for st in streams:
                st.synchronize()
        # the same bytes with one stream per direction (D2H of a chunk waits for its H2D through an event)
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def by_direction():
            for ci, c0 in enumerate(range(0, N, CH)):
                c1 = min(N, c0 + CH)
                with torch.cuda.stream(s_in):
                    d_in[offs[c0]:offs[c1]].copy_(tin[offs[c0]:offs[c1]], non_blocking=True)
                    d_off[c0:c1 + 1].copy_(t_off[c0:c1 + 1], non_blocking=True)
                    d_rc[c0:c1].copy_(t_rc[c0:c1], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(s_in)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev)
                    tout[c0 * 48:c1 * 48].copy_(d_out[c0 * 48:c1 * 48], non_blocking=True)
            s_in.synchronize(); s_out.synchronize()
        print("chunked copies, one stream per direction: %.2f ms" % timeit(by_direction), flush=True)
        print("chunked copies only (305 MB in, 48 MB out): %.2f ms free-running, %.2f ms with a sync at slot reuse" % (
            timeit(lambda: chunked(False)), timeit(lambda: chunked(True))), flush=True)
    # zero-copy experiment: k_pack16 reads the raw reads straight from pinned host memory (UVA pointer), no H2D of them
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    d_rc = torch.from_numpy(rc.astype(np.int32)).cuda()
    d_rec = torch.empty(N * 48, dtype=torch.uint8, device="cuda")
    t_rec = torch.from_numpy(h_rec.view(np.uint8).reshape(-1))
    st = torch.cuda.Stream()

    def zero_copy():
        with torch.cuda.stream(st):
            d_off.copy_(torch.from_numpy(h_off.view(np.int64)), non_blocking=True)
            d_rc.copy_(torch.from_numpy(h_rc.view(np.int32)), non_blocking=True)
            eng.align_batch_device(h_seqs.ctypes.data, d_off.data_ptr(), d_rc.data_ptr(), N, max_len, 0, d_rec.data_ptr(), 0, 0, st.cuda_stream)
            t_rec.copy_(d_rec, non_blocking=True)
        st.synchronize()
    print("zero-copy reads (one launch over the whole batch, records only): %.2f ms" % timeit(zero_copy), flush=True)
    for name, fl, ops in (("records+ops", cabi.WANT_OPS, h_ops), ("rec+gapped ops", cabi.WANT_OPS | cabi.OPS_GAPPED_ONLY, h_ops),
                          ("records only", 0, None)):
        ms = timeit(lambda: eng.align_batch(h_seqs, h_off, h_rc, fl, rec=h_rec, ops=ops))
        print("align_batch %-13s chunk=%s slots=%s: %.2f ms/step = %.1f M reads/s" % (
            name, os.environ.get("TREAT_B200_CHUNK", "default"), os.environ.get("TREAT_B200_SLOTS", "default"), ms, N / ms / 1e3), flush=True)


if __name__ == "__main__":
    main()
