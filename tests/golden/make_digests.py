#!/usr/bin/env python
"""Digests of the CPU oracle's results on the first 1,000,000 reads of bench.py's RPS12 workload
(BASELINE configs[2] / configs[3]): SHA-256 over the canonical 48-byte records, over the traceback ops and over the summary
counters.  bench.py's parity leg aligns the same reads sharded over N GPUs and must reproduce these at every N.

    python tests/golden/make_digests.py            # writes tests/golden/parity_digests.json (about a minute of CPU)
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import summary_from_records  # noqa: E402
from oracle import oracle as O  # noqa: E402
from treat_b200 import parallel, synth  # noqa: E402

WORKLOADS = {
    # name: (template, seed, p_trunc, reads) -- bench.py's make_template()
    "rps12": (synth.rps12_template, synth.SEED_BASE + 3, 0.0, 1_000_000),
}


def digests(name):
    make, seed, p_trunc, N = WORKLOADS[name]
    t = make()
    with tempfile.TemporaryDirectory() as d:
        om = O.NewTemplateFromFasta(t.write_fasta(os.path.join(d, "t.fa")), O.FORWARD, t.edit_base)
    seqs, off, rc = synth.reads(t, N, seed=seed, p_trunc=p_trunc)
    max_len = int(np.diff(off.astype(np.int64)).max())
    stride = ((t.m + max_len + 3) // 4 + 15) // 16 * 16
    rec, ops, _ = O.align_batch(om, seqs, off, rc, nthreads=os.cpu_count() or 1, ops_stride=stride)
    summary = summary_from_records(rec, t.m, 0, om.EditStop)
    return {"workload": name, "reads": N, "seed": seed, "p_trunc": p_trunc, "m": t.m, "K": t.K,
            "reads_sha256": parallel.sha256_hex(seqs, off, rc),
            "records_sha256": parallel.sha256_hex(parallel.canonical_records(rec)),
            "ops_sha256": parallel.sha256_hex(parallel.compact_ops(rec, ops)),
            "summary_sha256": parallel.sha256_hex(summary),
            "gapped_reads": int((rec["indel"] != 0).sum()), "reads_aligned": int(summary[14])}


if __name__ == "__main__":
    out = {k: digests(k) for k in WORKLOADS}
    path = os.path.join(ROOT, "tests", "golden", "parity_digests.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))
