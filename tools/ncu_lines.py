#!/usr/bin/env python
"""Per-CUDA-source-line instruction and sample shares from an .ncu-rep (needs -lineinfo + --import-source on).
Usage: python tools/ncu_lines.py rep.ncu-rep kernel_substr [reads_per_launch] [top_n]"""
import csv, io, subprocess, sys
rep, ksub = sys.argv[1], sys.argv[2]
reads = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
i = 0
agg = {}
src = {}
cur_fn = None
hdr = None
while i < len(rows):
    r = rows[i]
    if r and r[0] == "Function Name":
        cur_fn = r[1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and cur_fn and ksub in cur_fn and len(r) == len(hdr):
        ln = r[0]
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            ex, sm = int(r[ie] or 0), int(r[isamp] or 0)
        except ValueError:
            i += 1
            continue
        if r[2]:  # a SASS row (has an address)
            a = agg.setdefault(ln, [0, 0, 0])
            a[0] += ex; a[1] += sm; a[2] += 1
        if r[1]:
            src.setdefault(ln, r[1])
    i += 1
tot = sum(v[0] for v in agg.values()) or 1
tsm = sum(v[1] for v in agg.values()) or 1
print("total warp instr/read %.0f" % (tot / reads))
for ln, v in sorted(agg.items(), key=lambda x: -x[1][1])[:topn]:
    print("%5s  instr/read %7.1f (%4.1f%%)  samples %4.1f%%  sass %3d | %s" % (ln, v[0] / reads, 100 * v[0] / tot, 100 * v[1] / tsm, v[2], src.get(ln, "").strip()[:100]))
