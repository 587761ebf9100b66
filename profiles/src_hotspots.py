#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line:
warp-instructions executed, stall samples, shared-memory wavefronts.  usage: src_hotspots.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = dict(zip(hdr[4:], r[4:]))
        def num(k):
            try:
                return float(d.get(k, "0").replace(",", "") or 0)
            except ValueError:
                return 0.0
        out.append((num("Instructions Executed"), num("Warp Stall Sampling (All Samples)"), num("L1 Wavefronts Shared"),
                    num("L1 Wavefronts Shared Excessive"), cur_file, r[0], r[1].strip()[:90]))
tot_i = sum(o[0] for o in out) or 1
tot_s = sum(o[1] for o in out) or 1
print(f"total warp instructions {tot_i:.0f}, stall samples {tot_s:.0f}")
print(f"{'inst%':>6} {'stall%':>6} {'smem_wf':>10} {'excess':>9}  file:line  source")
for o in sorted(out, reverse=True)[:top]:
    print(f"{100 * o[0] / tot_i:6.2f} {100 * o[1] / tot_s:6.2f} {o[2]:10.0f} {o[3]:9.0f}  {o[4]}:{o[5]}  {o[6]}")
