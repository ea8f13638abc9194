#!/usr/bin/env python
"""The handful of raw metrics that say what bounds a kernel: python tools/ncu_brief.py rep.ncu-rep [kernel_substr]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts.avg"]
for r in rows[2:]:
    rec = dict(zip(h, r))
    if len(sys.argv) > 2 and sys.argv[2] not in rec.get("Kernel Name", ""):
        continue
    print(rec.get("Kernel Name", "")[:80])
    for k in keys:
        for hk in h:
            if hk.endswith(k):
                print("  %-70s %s %s" % (hk, rec[hk], u[h.index(hk)]))
    for hk in h:
        if "issue_stalled" in hk and "per_issue_active" in hk and float(rec[hk] or 0) > 0.15:
            print("  stall %-50s %s" % (hk.split("issue_stalled_")[1].split("_per_issue")[0], rec[hk]))
