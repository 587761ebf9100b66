#!/usr/bin/env python
"""Files the outputs of profiles/gpu_cycle.sh under profiles/: usage  collect.py <tag> [round]  (default round r2)."""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r2"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")
bench = os.path.join(src, f"{tag}_bench.json")
if os.path.exists(bench) and os.path.getsize(bench):
    shutil.copy(bench, os.path.join(dst, f"{rnd}_bench_{tag}_n1.json"))
    d = json.loads(open(bench).read().strip().splitlines()[-1])
    print(f"bench: {d['ms_per_step']:.3f} ms / step, e2e {d['e2e']['ms_per_step']:.3f} ms, {d['value'] / 1e9:.1f} G k-mers/s")
ll = os.path.join(src, f"{tag}_launches.csv")
if os.path.exists(ll) and os.path.getsize(ll):
    shutil.copy(ll, os.path.join(dst, f"{rnd}_launches_{tag}.csv"))
    agg = subprocess.run([sys.executable, os.path.join(dst, "launch_agg.py"), ll], capture_output=True, text=True).stdout
    open(os.path.join(dst, f"{rnd}_launches_{tag}_per_kernel.txt"), "w").write(agg)
    print(agg.split("\n")[0])
rep = os.path.join(src, f"{tag}_full.ncu-rep")
if os.path.exists(rep):
    raw = os.path.join(src, f"{tag}_full_raw.csv")
    open(raw, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
    subprocess.run([sys.executable, os.path.join(dst, "ncu_summary.py"), raw, os.path.join(dst, f"{rnd}_kernels_ncu_{tag}.json"),
                    f"profiles/gpu_cycle.sh {tag} <regex>", tag])
log = os.path.join(src, f"{tag}_pytest.log")
if os.path.exists(log):
    print("pytest:", open(log).read().strip().splitlines()[-1])
