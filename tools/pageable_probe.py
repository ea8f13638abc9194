"""treat_b200_align_batch from pageable host buffers (what a cgo caller passing Go slices has) against pinned ones."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, tempfile
import bench
import treat_b200 as T
from treat_b200 import cabi, synth
t = synth.rps12_template()
with tempfile.TemporaryDirectory() as d:
    tm = T.NewTemplateFromFasta(t.write_fasta(os.path.join(d, "t.fa")), T.FORWARD, "T")
eng = T.Aligner(0); eng.SetTemplate(tm)
N = 1000000
seqs, off, rc = synth.reads(t, N, seed=synth.SEED_BASE + 3)
flags = cabi.WANT_OPS | cabi.OPS_GAPPED_ONLY
stride = eng.ops_stride(int(np.diff(off.astype(np.int64)).max()))
def run(tag, seqs, off, rc, rec, ops):
    for it in range(4):
        t0 = time.perf_counter()
        eng.align_batch(seqs, off, rc, flags, rec=rec, ops=ops)
        dt = time.perf_counter() - t0
    print("%s: %.2f ms per 1M reads (%.1f M reads/s)" % (tag, 1e3 * dt, N / dt / 1e6), flush=True)
rec = np.zeros(N, dtype=cabi.RECORD_DTYPE); ops = np.zeros((N, stride), dtype=np.uint8)
run("pageable in, pageable out", seqs, off, rc, rec, ops)
hold = []
def pin(a):
    arr, p = bench.pinned_array(a.nbytes); hold.append(p)
    v = arr.view(a.dtype).reshape(a.shape); v[...] = a
    return v
pseqs, poff, prc = pin(seqs), pin(off), pin(rc)
run("pinned in, pageable out", pseqs, poff, prc, rec, ops)
prec, pops = pin(rec), pin(ops)
run("pinned in, pinned out", pseqs, poff, prc, prec, pops)
