"""treat_b200_fasta_stream / treat_b200_fasta_text from pageable file bytes against pinned ones."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, tempfile
import bench
import treat_b200 as T
from treat_b200 import cabi, synth
t = synth.rps12_template()
with tempfile.TemporaryDirectory() as d:
    tm = T.NewTemplateFromFasta(t.write_fasta(os.path.join(d, "t.fa")), T.FORWARD, "T")
eng = T.Aligner(0); eng.SetTemplate(tm)
N = 400000
seqs, off, rc = synth.reads(t, N, seed=5)
raw = seqs.tobytes()
fasta = b"".join(b">r%d-%d\n%s\n" % (i, rc[i], raw[int(off[i]):int(off[i + 1])]) for i in range(N))
pageable = np.frombuffer(fasta, dtype=np.uint8)
pinned, hold = bench.pinned_array(len(fasta)); pinned[:] = pageable
cnt = [0]
scb = cabi.FASTA_SINK(lambda u, pp: 0)
tcb = cabi.TEXT_SINK(lambda u, p, n, f, k: 0)
tot = C.c_uint64()
for tag, buf in (("pageable", pageable), ("pinned", pinned)):
    for name, fn in (("fasta_stream", lambda: cabi.lib().treat_b200_fasta_stream(eng._ctx, buf.ctypes.data, buf.size, cabi.WANT_OPS | cabi.OPS_GAPPED_ONLY, 0, scb, None, C.byref(tot))),
                     ("fasta_upload", lambda: cabi.lib().treat_b200_fasta_upload(eng._ctx, buf.ctypes.data, buf.size, C.byref(tot), None)),
                     ("fasta_text", lambda: cabi.lib().treat_b200_fasta_text(eng._ctx, buf.ctypes.data, buf.size, cabi.NO_SUMMARY, 80, tcb, None, C.byref(tot)))):
        for it in range(3):
            t0 = time.perf_counter()
            cabi.check(fn(), eng._ctx)
            dt = time.perf_counter() - t0
        print("%s %s: %.2f ms per %d reads" % (name, tag, 1e3 * dt, N), flush=True)
