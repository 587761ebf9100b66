#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel over ONE resident step of bench.py
(the launches between two consecutive read_sketch_kernel launches).  usage: launch_agg.py launches.csv [step_index]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr = r
        rows = rows[i + 1:]
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
L = [(r[ki].split("(")[0].replace("void ", ""), float(r[vi].replace(",", ""))) for r in rows]
idx = [i for i, (n, t) in enumerate(L) if "read_sketch_kernel" in n]
steps = [(s, e) for s, e in zip(idx, idx[1:] + [len(L)]) if e - s > 8]
s, e = steps[int(sys.argv[2]) if len(sys.argv) > 2 else 2]
seg = L[s:e]
agg = collections.OrderedDict()
for n, t in seg:
    agg.setdefault(n, [0, 0])
    agg[n][0] += t
    agg[n][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"step launches {len(seg)}, serialized kernel time {tot / 1000:.1f} us")
for n, (t, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{t / 1000:9.1f} us {c:3d} {100 * t / tot:5.1f}% {n}")
