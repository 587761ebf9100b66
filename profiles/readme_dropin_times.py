#!/usr/bin/env python
"""BASELINE configs[0] through the drop-in binary: oracle/_ref/PHI_gpu (reference CLI + libphi_gpu_index.so) against oracle/_ref/PHI_ref
(unmodified reference), same README inputs (tests/golden/mhc4.npz written back to GFA / FASTA), front-end wall = difference of the
reference's own log stamps 'Graph has' -> 'Filtered/Retained' (/root/reference/src/ILP_index.cpp:537,738).  Also the real-data
walk-sharing figures through the C ABI.  usage (GPU box): readme_dropin_times.py > out.json"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from golden_cases import Case
from phi_b200 import synth
import phi_b200


def stamps(exe, gfa, fa, tmp, env=None, threads=None):
    e = dict(os.environ, PHI_STUB_DUMP=os.path.join(tmp, "dump.txt"), PHI_ADAPTER_TIMES="1")
    e.update(env or {})
    t0 = time.time()
    p = subprocess.Popen([exe, "-g", gfa, "-r", fa, "-o", os.path.join(tmp, "o.fa"), "-t", str(threads or os.cpu_count())], env=e,
                         stderr=subprocess.PIPE, stdout=subprocess.DEVNULL, text=True)
    st = {}
    for line in p.stderr:
        if line.startswith("[phi_adapter]"):
            st["adapter"] = line.strip()                   # flat views / wait for the CUDA context / run
        m = re.match(r"\[M::ILP_function::([\d.]+)\*", line)
        if m:
            for key in ("Graph has", "Haplotypes sketched", "Indexed reads", "Filtered/Retained"):
                if key in line:
                    st[key] = float(m.group(1))
        if "Filtered/Retained" in line:
            st["wall_to_filtered_s"] = time.time() - t0
            break
    p.kill(); p.wait()
    return st


c = Case("mhc4")
out = {"config": "BASELINE configs[0]: README test (MHC_4.gfa.gz + CHM13_reads.fq.gz), -k31 -w25", "host_threads": os.cpu_count()}
try:
    out["gpu_persistence_mode"] = subprocess.run(["nvidia-smi", "--query-gpu=persistence_mode", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    # what a fresh process pays before it can launch anything: phi_gpu_index_create = CUDA context creation + a few small allocations
    code = ("import ctypes, time; lib = ctypes.CDLL(%r); ctx = ctypes.c_void_p(); t = time.time(); "
            "rc = lib.phi_gpu_index_create(0, ctypes.byref(ctx)); print(rc, time.time() - t)") % os.path.join(ROOT, "phi_b200", "libphi_gpu_index.so")
    out["ctx_create_s_fresh_process"] = [float(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout.split()[1]) for _ in range(3)]
except Exception:
    pass
with tempfile.TemporaryDirectory() as tmp:
    gfa, fa = os.path.join(tmp, "g.gfa"), os.path.join(tmp, "r.fa")
    synth.write_gfa(c.graph, gfa)
    synth.write_fasta(c.reads, fa)
    for name in ("PHI_ref", "PHI_gpu", "PHI_gpu_model"):
        exe = os.path.join(ROOT, "oracle", "_ref", name)
        runs = [stamps(exe, gfa, fa, tmp) for _ in range(3 if name != "PHI_ref" else 1)]
        best = min(runs, key=lambda s: s["Filtered/Retained"] - s["Graph has"])
        out[name] = {"front_end_wall_s": round(best["Filtered/Retained"] - best["Graph has"], 4), "stamps": best,
                     "all_front_end_walls_s": [round(s["Filtered/Retained"] - s["Graph has"], 4) for s in runs]}
ix = phi_b200.PhiGpuIndex(0)
ix.upload(c.graph, c.reads)
for share in (True, False):
    ix.set_walk_sharing(11, share)
    for _ in range(3):
        r = ix.run_resident(c.k, c.w, c.T, download=False)
    ts = []
    for _ in range(10):
        r = ix.run_resident(c.k, c.w, c.T, download=False)
        ts.append(ix.times())
    sh = ix.sharing()
    out["resident_share%d" % share] = {"ms_per_step": float(np.mean([t["total_ms"] for t in ts])), "walk_kernel_ms": float(np.mean([t["walk_kernel_ms"] for t in ts])),
                                       "unique_windows": sh["unique_windows"], "path_kmer_positions": r.path_kmer_positions,
                                       "unique_fraction": sh["unique_windows"] / r.path_kmer_positions}
ix.close()
out["speedup_front_end"] = out["PHI_ref"]["front_end_wall_s"] / out["PHI_gpu"]["front_end_wall_s"]
print(json.dumps(out, indent=1))
