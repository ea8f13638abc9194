import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, "/root/repo")
import tempfile, bench
import treat_b200 as T
from treat_b200 import cabi, synth
N = 1000000
t = synth.rps12_template()
with tempfile.TemporaryDirectory() as d:
    tm = T.NewTemplateFromFasta(t.write_fasta(os.path.join(d, "tmpl.fa")), T.FORWARD, t.edit_base)
eng = T.Aligner(0); eng.SetTemplate(tm)
seqs, off, rc = synth.reads(t, N, seed=synth.SEED_BASE + 3)
parts = []
for i in range(N):
    parts.append(b">r%d_%d\n" % (i, rc[i])); parts.append(seqs[int(off[i]):int(off[i + 1])].tobytes()); parts.append(b"\n")
text = b"".join(parts)
arr, p = bench.pinned_array(len(text)); arr[:] = np.frombuffer(text, dtype=np.uint8)
seen = [0]
def sink(_u, pp):
    seen[0] += pp.contents.n
    return 0
cb = cabi.FASTA_SINK(sink); tot = C.c_uint64()
flags = cabi.WANT_OPS | cabi.OPS_GAPPED_ONLY
for it in range(5):
    t0 = time.perf_counter()
    cabi.check(cabi.lib().treat_b200_fasta_stream(eng._ctx, arr.ctypes.data, arr.size, flags, 0, cb, None, C.byref(tot)), eng._ctx)
    print("stream %.2f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
