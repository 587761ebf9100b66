#!/usr/bin/env python
"""Experiment: resident step time of one bench config as a function of the walk-chunk bucket size (phi_gpu_index_set_walk_sharing).
usage: exp_sweep.py <config> [shift ...]   (GPU box; prints one JSON line per shift)"""
import json
import sys
import os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cfg = sys.argv[1]
shifts = [int(x) for x in sys.argv[2:]] or [10, 11, 12, 13, 14]
sys.argv = ["bench.py", "--config", cfg]
import bench
import phi_b200
a = bench.parse_args()
wl = bench.Workload(a)
g, rd, base, nw, region, units = wl.shard(0, 1)
ix = phi_b200.PhiGpuIndex(0)
ix.upload(g, rd)
for sh in shifts:
    ix.set_walk_sharing(sh, True)
    for _ in range(2):
        ix.run_resident(a.k, a.w, 1.0, download=False)
    ts = []
    for _ in range(4):
        ix.run_resident(a.k, a.w, 1.0, download=False)
        ts.append(ix.times())
    tm = {k: round(float(np.mean([t[k] for t in ts])), 3) for k in ts[0]}
    print(json.dumps({"config": cfg, "chunk_shift": sh, "times": tm, "sharing": ix.sharing()}), flush=True)
ix.close()
