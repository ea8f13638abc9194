#!/usr/bin/env python
"""Warp instructions per opcode (and shared-memory wavefronts) of one kernel in an .ncu-rep captured with --import-source on:
python tools/ncu_ops.py rep.ncu-rep reads_per_launch [warp_iterations]"""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
reads = float(sys.argv[2]); iters = float(sys.argv[3]) if len(sys.argv) > 3 else reads
rows = list(csv.reader(io.StringIO(out)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[start]; ix = {k: i for i, k in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[start + 1:]:
    if len(r) != len(hdr): continue
    src = r[ix["Source"]].split()
    op = src[1] if src[0].startswith("@") else src[0]
    try:
        ex = int(r[ix["Instructions Executed"]] or 0); wf = int(r[ix["L1 Wavefronts Shared"]] or 0); wfi = int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    except ValueError: continue
    parts = op.split(".")
    key = parts[0] + ("." + ".".join(parts[1:3]) if parts[0] in ("LDS", "STS", "ATOMS", "LDGSTS", "IMAD") else "")
    a = agg[key]; a[0] += ex; a[1] += wf; a[2] += wfi
tot = sum(v[0] for v in agg.values())
print("total %.1f warp instr/read" % (tot / reads))
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:45]:
    print("%-22s instr/read %6.2f  shared wavefronts/iter %7.1f (ideal %7.1f)" % (k, v[0] / reads, v[1] / iters, v[2] / iters))
