import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, tempfile
import treat_b200 as T
from treat_b200 import cabi, synth
t = synth.rps12_template()
with tempfile.TemporaryDirectory() as d:
    tm = T.NewTemplateFromFasta(t.write_fasta(os.path.join(d, "t.fa")), T.FORWARD, "T")
eng = T.Aligner(0); eng.SetTemplate(tm)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
seqs, off, rc = synth.reads(t, N, seed=5)
raw = seqs.tobytes()
fasta = b"".join(b">r%d-%d\n%s\n" % (i, rc[i], raw[int(off[i]):int(off[i+1])]) for i in range(N))
import ctypes as C
tot = [0]
def sink(u, p, n, f, k):
    tot[0] += n; return 0
cb = cabi.TEXT_SINK(sink)
nrec = C.c_uint64()
ptr = C.c_void_p()
cabi.check(cabi.lib().treat_b200_host_alloc(C.byref(ptr), len(fasta)))  # pinned, as bench.py's buffers
buf = np.frombuffer((C.c_uint8 * len(fasta)).from_address(ptr.value), dtype=np.uint8)
buf[:] = np.frombuffer(fasta, dtype=np.uint8)
for rep in range(3):
    tot[0] = 0
    t0 = time.perf_counter()
    cabi.check(cabi.lib().treat_b200_fasta_text(eng._ctx, buf.ctypes.data, buf.size, cabi.NO_SUMMARY, 80, cb, None, C.byref(nrec)), eng._ctx)
    print("rep", rep, nrec.value, tot[0], "%.2f ms" % (1e3 * (time.perf_counter() - t0)))
