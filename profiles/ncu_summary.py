#!/usr/bin/env python
"""Summarise an `ncu --set full` report (exported with `ncu -i rep --page raw --csv`) into a small JSON: one entry per kernel
with the metrics the roofline discussion in DESIGN.md uses.  usage: ncu_summary.py raw.csv out.json "<command>" "<version note>" """
import csv
import json
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
out = {"command": sys.argv[3], "version": sys.argv[4], "kernels": []}
for r in rows[2:]:
    k = {"name": r[hdr.index("Kernel Name")][:60]}
    for w in WANT:
        if w in hdr:
            k[w] = [r[hdr.index(w)], units[hdr.index(w)]]
    out["kernels"].append(k)
json.dump(out, open(sys.argv[2], "w"), indent=1)
for k in out["kernels"]:
    print(k["name"], k.get("gpu__time_duration.sum"), "issue", k.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
          "dram", k.get("dram__bytes_read.sum"), k.get("dram__bytes_write.sum"))
