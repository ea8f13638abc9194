#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): per-kernel headline metrics + the hottest SASS
regions.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [reads_per_launch] > profiles/x.md"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
reads = float(sys.argv[2]) if len(sys.argv) > 2 else None
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_integer_pred_on.sum",
]
STALLS = ["short_scoreboard", "long_scoreboard", "wait", "math_pipe_throttle", "not_selected", "selected",
          "branch_resolving", "barrier", "dispatch_stall", "mio_throttle", "no_instructions", "lg_throttle"]


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units = raw[0], raw[1]
print("# ncu summary of `%s`\n" % rep)
print("Captured with `ncu --set full --clock-control none --import-source on` under gpurun (B200, sm_100a).\n")
for r in raw[2:]:
    print("## %s\n" % r[hdr.index("Kernel Name")])
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in hdr:
            print("| %s | %s | %s |" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    tot = 0
    st = {}
    for s in STALLS:
        k = "smsp__pcsamp_warps_issue_stalled_" + s
        if k in hdr:
            st[s] = float(r[hdr.index(k)] or 0)
            tot += st[s]
    if tot:
        print("\nWarp-state samples: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(st.items(), key=lambda x: -x[1]) if v / tot > 0.005))
    if reads and "dram__bytes_read.sum" in hdr:
        def tobytes(v, u):
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        rd = tobytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = tobytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        print("\nDRAM traffic per read: %.0f B read + %.0f B written (%.0f reads per launch)" % (rd / reads, wr / reads, reads))
        if "smsp__inst_executed.sum" in hdr:
            print("\nWarp instructions per read: %.0f" % (float(r[hdr.index("smsp__inst_executed.sum")]) / reads))
    print()

src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv"))))
starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
seen = set()
for si, st0 in enumerate(starts):
    name = src[st0][1]
    if name in seen:
        continue
    seen.add(name)
    sec = src[st0:(starts[si + 1] if si + 1 < len(starts) else len(src))]
    h = sec[1]
    data = [r for r in sec[2:] if len(r) == len(h)]
    ie, isamp, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    tot = sum(int(r[ie]) for r in data) or 1
    tsamp = sum(int(r[isamp]) for r in data) or 1
    print("## SASS hot regions: %s\n" % name)
    print("| SASS lines | share of warp instructions | share of samples | dominant opcodes |\n|---|---|---|---|")
    for s in range(0, len(data), 50):
        ch = data[s:s + 50]
        ex = sum(int(r[ie]) for r in ch)
        sm = sum(int(r[isamp]) for r in ch)
        if ex / tot > 0.02:
            ops = Counter((r[isrc].split()[1] if r[isrc].strip().startswith("@") else r[isrc].split()[0]) for r in ch)
            print("| %d-%d | %.1f%% | %.1f%% | %s |" % (s, s + 49, 100 * ex / tot, 100 * sm / tsamp,
                                                  ", ".join("%s x%d" % kv for kv in ops.most_common(5))))
    print()
