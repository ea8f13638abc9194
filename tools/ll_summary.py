#!/usr/bin/env python
"""Average duration per kernel of an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/ll_summary.py file.csv [skip_first_n_per_kernel]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < len(h):
        continue
    rec = dict(zip(h, r))
    if rec.get("Metric Name") == "gpu__time_duration.sum":
        d.setdefault(rec["Kernel Name"].split("(")[0][:50], []).append(float(rec["Metric Value"]) / (1000.0 if rec["Metric Unit"] == "ns" else 1.0))
tot = 0.0
for k, v in d.items():
    w = v[skip:] or v
    tot += sum(w) / len(w)
    print("%-52s n=%3d  %8.1f us" % (k, len(v), sum(w) / len(w)))
print("sum of averages: %.1f us" % tot)
