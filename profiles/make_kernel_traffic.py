#!/usr/bin/env python
"""profiles/kernel_traffic.json (what bench.py reports as roofline.traffic / issue_slot_util / warp_inst_per_kmer) from ncu summaries.
usage: make_kernel_traffic.py <config>=<ncu summary json>:<bench line json> [...]
The ncu summary is written by ncu_summary.py from `ncu --set full --clock-control none` of `bench.py --config <config> --steps 2
--warmup 3 --no-cpu-baseline --no-extras`; the bench line of the same command gives the units one launch processes."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
if "walk_sketch_kernel" in out:
    out = {}                                             # round-1 layout (not keyed by config)
for arg in sys.argv[1:]:
    cfg, rest = arg.split("=", 1)
    summ_ps, bench_p = rest.split(":", 1)
    bench = json.loads([l for l in open(bench_p) if l.startswith("{")][-1])
    units = {"walk_sketch_kernel": bench["sharing"]["unique_windows"], "read_sketch_kernel": bench["units_per_step"]["read_kmer_positions"]}
    summs = [json.load(open(q)) for q in summ_ps.split("+")]          # later summaries override earlier ones kernel by kernel
    ent = {"source": "ncu --set full --clock-control none, " + " + ".join(f"profiles/{os.path.basename(q)}" for q in summ_ps.split("+")) + f" ({summs[-1]['version']})"}
    for k in [k for sm in summs for k in sm["kernels"]]:
        name = k["name"].split("(")[0].split("<")[0].replace("void ", "").replace("phi::", "")
        for suffix in ("_r64", "_r72"):                      # register-budget variants of the walk kernel
            if name.endswith(suffix):
                name = name[:-len(suffix)]
        def val(m):
            v = k.get(m)
            if not v:
                return None
            x = float(v[0].replace(",", ""))
            u = v[1]
            return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        e = {"dram_bytes_per_launch": (rd or 0) + (wr or 0), "dram_read": rd, "dram_write": wr,
             "gpu_time_us": val("gpu__time_duration.sum"), "issue_slot_util": (val("smsp__issue_active.avg.pct_of_peak_sustained_active") or 0) / 100.0,
             "alu_pipe_util": (val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") or 0) / 100.0,
             "warps_active": (val("sm__warps_active.avg.pct_of_peak_sustained_active") or 0) / 100.0,
             "warp_inst": val("smsp__inst_executed.sum"), "registers": val("launch__registers_per_thread")}
        if name in units and e["warp_inst"]:
            e["warp_inst_per_kmer"] = e["warp_inst"] / units[name]
        ent[name] = e
    out[cfg] = ent
json.dump(out, open(out_path, "w"), indent=1)
print("wrote", out_path, list(out))
