#!/usr/bin/env python3
"""Condenses an `ncu --page raw --csv` export into a small JSON (one record per launch, selected metrics) for profiles/.
usage: tools/ncu_summary.py <raw.csv> <out.json> "<command line that was profiled>" """
import csv
import json
import sys

WANT = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, units = rows[0], rows[1]
out = {"command": sys.argv[3] if len(sys.argv) > 3 else "", "launches": []}
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    rec = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            rec[w] = (r[i] + (" " + units[i] if units[i] and w not in ("Kernel Name", "Grid Size", "Block Size") else "")).strip()
    out["launches"].append(rec)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(len(out["launches"]), "launches")
