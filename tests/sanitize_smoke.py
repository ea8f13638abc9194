#!/usr/bin/env python
"""Small run of every kernel variant for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tests/sanitize_smoke.py
Checks results against the oracle as it goes (test tooling)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import treat_b200 as T  # noqa: E402
from treat_b200 import cabi, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def check(eng, t, N, seed, p_trunc, flags=0):
    d = tempfile.mkdtemp()
    path = t.write_fasta(os.path.join(d, "t.fa"))
    tm, om = T.NewTemplateFromFasta(path, T.FORWARD, "T"), O.NewTemplateFromFasta(path, O.FORWARD, "T")
    eng.SetTemplate(tm)
    seqs, off, rc = synth.reads(t, N, seed=seed, p_trunc=p_trunc)
    rec, ops = eng.align_batch(seqs, off, rc, cabi.WANT_OPS | flags)
    orec, oops, _ = O.align_batch(om, seqs, off, rc, nthreads=4, ops_stride=ops.shape[1])
    for f in ("edit_stop", "junc_start", "junc_end", "junc_len", "score", "mismatches", "indel", "alt_editing", "aln_len"):
        assert np.array_equal(rec[f], orec[f]), f
    assert np.array_equal(ops, oops)


eng = T.Aligner(0)
rps, lng = synth.rps12_template(), synth.long_template(300, 8)
check(eng, rps, 1501, 1, 0.0)
check(eng, rps, 601, 2, 0.5)                                  # band fallback -> global scratch
check(eng, rps, 301, 3, 0.2, cabi.ONE_READ_PER_WARP)
check(eng, rps, 301, 4, 0.2, cabi.FULL_TRACEBACK_STORE)
check(eng, lng, 401, 5, 0.2)                                  # u32 direction words, K = 10
check(eng, synth.long_template(40, 2), 301, 6, 0.3)
check(eng, rps, 20001, 7, 0.05, cabi.DP_EVERY_READ)           # the paired DP kernel on every read
check(eng, synth.long_template(700, 4), 201, 8, 0.2)          # column strips (m > 480)
check(eng, synth.long_template(230, 3), 2001, 9, 0.1)         # fused kernel, one read per warp (257..512 edit sites)
# wide path: template with a non-alphabet base, read with a saturated edit-base run
d = tempfile.mkdtemp()
p = os.path.join(d, "w.fa")
open(p, "w").write(">fe\nACNGTTAGC%s\n>pe\nACNGAGC%s\n" % ("AG" * 20, "AG" * 20))
tm = T.NewTemplateFromFasta(p, T.FORWARD, "T")
alns = eng.NewAlignments([("a-1", "ACNGTTAGC" + "AG" * 20), ("b-2", "ACNG" + "T" * 300 + "AGC" + "AG" * 20)], tm, False)
assert alns[0].Score == 47
# bounds-check flags of the kernels' data-dependent index computations (a build with -DTREAT_B200_CHECKS has them compiled in)
flags, compiled = eng.debug_flags()
print("debug_flags 0x%x checks_compiled %d" % (flags, compiled))
assert flags == 0, "a kernel index left its bounds: 0x%x" % flags
if os.environ.get("TREAT_B200_EXPECT_CHECKS"):
    assert compiled, "this library was built without TREAT_B200_CHECKS"
print("sanitize smoke ok")
eng.close()
