#!/usr/bin/env python
"""bench.py — the ILP_index front end on N B200s, one JSON line (driver contract).

A "step" is one pass of the hot path (/root/reference/src/ILP_index.cpp:543-743: read sketch ->
ranked spectrum -> walk sketch + match -> threshold filter -> anchor CSR) over one batch of synthetic
input of the shape BASELINE.json configs[1] names: an MHC-shaped acyclic graph, 49 haplotypes x ~5 Mbp,
150 bp reads at 10x.

  value  : (read k-mer positions + path k-mer positions) / s, inputs resident in HBM, wall clock over K steps
           bracketed by device synchronisation (the pipeline's own CUDA-event stage times are reported too).
  e2e    : the same through phi_gpu_index_run() — host buffers in, host CSR out, H2D/D2H inside the timed region.
  --impl reference : the UNMODIFIED reference CLI (oracle/_ref/PHI_ref: reference sources + recording Gurobi
           stub) on the box's host cores, bounded sample of the same workload, front-end stage timed from
           the reference's own log stamps.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x50484931 + 1          # "PHI1" + config index
METRIC = "read k-mers counted/s + path k-mers matched/s (ILP_index front end)"
UNIT = "k-mers/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="phi_b200", choices=["phi_b200", "reference"])
    ap.add_argument("--haps", type=int, default=49)
    ap.add_argument("--backbone", type=int, default=5_000_000)
    ap.add_argument("--coverage", type=float, default=10.0)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("-w", type=int, default=25)
    ap.add_argument("--cpu-walks", type=int, default=0, help="walks in the CPU-baseline sample (0: min(nproc, haps, 16))")
    ap.add_argument("--cpu-coverage", type=float, default=0.5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(args):
    from phi_b200 import synth
    sg = synth.make_graph(SEED, args.backbone, args.haps)
    rd = synth.make_reads(SEED, sg, args.coverage, read_len=args.read_len)
    return sg, rd


def positions(lengths, k, w):
    lengths = np.asarray(lengths, dtype=np.int64)
    return int(np.sum(np.where(lengths >= w + k - 1, lengths - k + 1, 0)))


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU reference (bounded sample)
def run_reference_sample(sg, rd, args, n_walks, cov, threads):
    """Times the reference's own front end (log-stamp deltas) on `n_walks` walks of the graph and a
    `cov`-coverage prefix of the reads.  Kills the process once the front end is done (model construction is not timed)."""
    from phi_b200 import synth
    exe = os.path.join(ROOT, "oracle", "_ref", "PHI_ref")
    g = sg.graph
    n_reads = max(1, int(rd.n_reads * cov / args.coverage))
    sub_reads = rd.take(0, n_reads)
    wl = g.walk_lengths()[:n_walks]
    P = positions(wl, args.k, args.w)
    Q = positions(np.diff(sub_reads.read_off.astype(np.int64)), args.k, args.w)
    if not os.path.exists(exe):
        # the reference did not compile here: time the plain-C port instead (kind "port")
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import phi_io
        t0 = time.time()
        phi_io.oracle_index(g.take_walks(0, n_walks), sub_reads, args.k, args.w, 1.0, threads)
        dt = time.time() - t0
        return dict(kind="port", P=P, Q=Q, t_index=dt, t_reads=None, t_paths=None)
    with tempfile.TemporaryDirectory() as tmp:
        gfa, fa = os.path.join(tmp, "g.gfa"), os.path.join(tmp, "r.fa")
        synth.write_gfa(g, gfa, walks=range(n_walks))
        synth.write_fasta(sub_reads, fa)
        env = dict(os.environ, PHI_STUB_DUMP=os.path.join(tmp, "dump.txt"))
        p = subprocess.Popen([exe, "-g", gfa, "-r", fa, "-o", os.path.join(tmp, "o.fa"), "-t", str(threads), "-k", str(args.k),
                              "-w", str(args.w)], env=env, stderr=subprocess.PIPE, stdout=subprocess.DEVNULL, text=True)
        stamps = {}
        for line in p.stderr:
            m = re.match(r"\[M::ILP_function::([\d.]+)\*", line)
            if m:
                for key in ("Graph has", "Haplotypes sketched", "Indexed reads", "Filtered/Retained"):
                    if key in line:
                        stamps[key] = float(m.group(1))
            if "Filtered/Retained" in line:
                break
        p.kill()
        p.wait()
    t_a = stamps["Haplotypes sketched"] - stamps["Graph has"]
    t_b = stamps["Indexed reads"] - stamps["Haplotypes sketched"]
    t_c = stamps["Filtered/Retained"] - stamps["Indexed reads"]
    return dict(kind="reference", P=P, Q=Q, t_index=t_a + t_b + t_c, t_reads=t_b, t_paths=t_a + t_c)


def cpu_sample_desc(n_walks, cov, args):
    return (f"{n_walks} of {args.haps} walks (~{args.backbone / 1e6:g} Mbp each) + {cov:g}x of the {args.read_len} bp reads; "
            f"front-end stage only (log stamps 'Graph has' -> 'Filtered/Retained')")


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sg, rd = workload(args)
    n_walks = args.cpu_walks or min(threads, args.haps, 16)
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        r = run_reference_sample(sg, rd, args, n_walks, args.cpu_coverage, threads)
        if i >= args.warmup:
            vals.append((r["P"] + r["Q"]) / r["t_index"])
            times.append(r["t_index"])
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(times)) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u64", "data": "synthetic",
            "config": config_dict(args, 1),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": r["kind"],
                             "sample": cpu_sample_desc(n_walks, args.cpu_coverage, args)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def config_dict(args, n):
    return {"workload": f"BASELINE configs[1]: synthetic MHC-shaped acyclic graph, {args.haps} haplotypes x ~{args.backbone / 1e6:g} Mbp, "
                        f"nodes chopped to <=30 bp, {args.read_len} bp reads at {args.coverage:g}x, k={args.k} w={args.w} T=1.0",
            "seed": SEED, "gpus": n, "cpu_affinity": getattr(args, "cpu_affinity", "unchanged"),
            "l2": "no explicit flush: per-step working set (walk steps, step offsets, chunk table, reads, "
                  "spectrum table, hit and anchor buffers: > 400 MB) exceeds the 126 MB L2"}


def bind_to_gpu_numa(device):
    """Pin this process (and the threads / pinned allocations that follow) to the CPUs of the NUMA node the GPU hangs off: host <-> device
    copies of the end-to-end arm then stay on the local socket.  Plumbing only; silently does nothing where sysfs / NVML say nothing."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:                      # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        cpus = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return f"numa-local ({len(ids)} cpus: {cpus})"
    except Exception:
        pass
    return "unchanged"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def walk_kernel_algorithmic_bytes(res, g, k):
    """SURVEY.md §8(d) per-unit figure x units of one launch, with this run's measured densities:
    0.25 B/position 2-bit segment store + 4 B per walk step (vertex id) + 32 B probe sector per emitted minimizer
    + per hit a 16 B record and 4 B per anchor vertex.  Units = ALL path k-mer positions the launch accounts for (the
    result covers every walk); with walk sharing the kernel physically sketches only the representative chunks, see
    roofline.sharing."""
    P = res.path_kmer_positions
    steps = len(g.walk_vtx)
    hit_vtx = res.path_hits * (1.0 + (k - 1) / max(1.0, (g.walk_lengths().sum() / max(1, steps))))
    return 0.25 * P + 4.0 * steps + 32.0 * res.path_minimizers_emitted + 16.0 * res.path_hits + 4.0 * hit_vtx


def read_kernel_algorithmic_bytes(res, rd):
    """SURVEY.md §8(d), the part of the per-read-k-mer figure that belongs to the read sketch kernel: the ASCII bases once
    + per emitted minimizer a 32 B table sector read and written."""
    return float(rd.read_bases.nbytes) + 64.0 * res.read_minimizers_emitted


def kernel_traffic(name):
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    try:
        return float(json.load(open(tp))[name]["dram_bytes_per_launch"])
    except Exception:
        return None


def main_gpu(args):
    import phi_b200
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    all_cpus = os.sched_getaffinity(0)
    args.cpu_affinity = bind_to_gpu_numa(local)
    if world > 1:
        from phi_b200 import multi
        return multi.bench_main(args, rank, world, local, globals())
    sg, rd = workload(args)
    g = sg.graph
    ix = phi_b200.PhiGpuIndex(local)
    k, w = args.k, args.w

    # ---- resident: inputs already in HBM when the timed region starts
    ix.upload(g, rd)
    clocks = ClockSampler(local)                                # samples from the first warm-up step to the end of the e2e loop:
    clocks.start()                                              # the GPU is busy for that whole interval (steps are ~10 ms each)
    for _ in range(args.warmup):
        res = ix.run_resident(k, w, 1.0, download=False)
    stage = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = ix.run_resident(k, w, 1.0, download=False)        # ends with a stream synchronise
        stage.append(ix.times())
    dt = time.perf_counter() - t0
    units = res.read_kmer_positions + res.path_kmer_positions
    value = units * args.steps / dt
    tm = {key: float(np.mean([s[key] for s in stage])) for key in stage[0]}

    # ---- end to end through the drop-in call: host buffers -> host CSR
    # inputs sit in pinned host memory (phi_gpu_host_alloc), the result lands in pinned memory owned by the library;
    # every step copies all inputs host->device and the whole CSR device->host inside the timed region
    gp, rp = ix.pinned_inputs(g, rd)
    for _ in range(args.warmup):
        ix.free_raw(ix.run_raw(gp, rp, k, w, 1.0))
    e2e_stage = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        raw = ix.run_raw(gp, rp, k, w, 1.0)
        n_anchors_seen = raw.contents.n_anchors           # the caller reads the result
        ix.free_raw(raw)
        e2e_stage.append(ix.times())
    dt_e2e = time.perf_counter() - t0
    clk = clocks.stop()
    full = ix.run(gp, rp, k, w, 1.0)
    assert full.n_anchors == n_anchors_seen
    h2d = sum(a.nbytes for a in (g.seg_off, g.seg_bases, g.walk_off, g.walk_vtx, g.top_order_map, rd.read_off, rd.read_bases))
    d2h = full.wire_bytes()
    e2e_value = units * args.steps / dt_e2e

    peak, peak_src = measured_peak()
    share = ix.sharing()
    kernels = {}
    for name, ms, alg in (("walk_sketch_kernel", tm["walk_kernel_ms"], walk_kernel_algorithmic_bytes(res, g, k)),
                          ("read_sketch_kernel", tm["read_kernel_ms"], read_kernel_algorithmic_bytes(res, rd))):
        ach = alg / (ms * 1e-3) / 1e9
        kernels[name] = {"kernel_ms": ms, "algorithmic_bytes_per_launch": alg, "achieved": ach, "frac": ach / peak, "traffic": kernel_traffic(name)}
    dom = max(kernels, key=lambda n: kernels[n]["kernel_ms"])                # the dominant kernel of the step
    frac_unique = share["unique_windows"] / max(1, res.path_kmer_positions)
    kernels["walk_sketch_kernel"]["achieved_physical"] = kernels["walk_sketch_kernel"]["achieved"] * frac_unique

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u64", "data": "synthetic", "config": config_dict(args, 1),
            "read_kmers_per_s": res.read_kmer_positions / ((tm["read_sketch_ms"] + tm["spectrum_ms"]) * 1e-3),
            "path_kmers_per_s": res.path_kmer_positions / ((tm["graph_prep_ms"] + tm["walk_sketch_ms"] + tm["filter_ms"]) * 1e-3),
            "index_wall_s": dt_e2e / args.steps,
            "units_per_step": {"read_kmer_positions": res.read_kmer_positions, "path_kmer_positions": res.path_kmer_positions,
                               "read_minimizers": res.read_minimizers_emitted, "path_minimizers": res.path_minimizers_emitted,
                               "path_hits": res.path_hits, "spectrum": res.count_sp_r, "anchors": full.n_anchors,
                               "filtered_ranks": full.n_filtered},
            "stage_ms": tm,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": dt_e2e / args.steps * 1e3, "host_memory": "pinned (phi_gpu_host_alloc) in, pinned (library pool) out",
                    "stage_ms": {key: float(np.mean([x[key] for x in e2e_stage])) for key in e2e_stage[0]}},
            "gpu_launches": int(tm["kernel_launches"]) * args.steps,
            "clocks": clk,
            "device_ms_per_step": tm["total_ms"],
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved"], "peak": peak, "unit": "GB/s",
                         "frac": kernels[dom]["frac"], "traffic": kernels[dom]["traffic"], "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes_per_launch"], "kernel_ms": kernels[dom]["kernel_ms"],
                         "kernels": kernels,
                         "sharing": dict(share, path_kmer_positions=res.path_kmer_positions, unique_fraction=frac_unique),
                         "note": "both sketch kernels are integer-issue bound (~7 warp instructions per sketched k-mer, ALU pipe 60-72 % busy, vs a few bytes "
                                 "of compulsory traffic), so their HBM fraction is low by construction (ncu: DRAM traffic per launch in "
                                 "'traffic'). walk_sketch_kernel: 'achieved' counts the algorithmic bytes of ALL path k-mers the launch "
                                 "accounts for; identical walk chunks are sketched once (sharing.unique_fraction), 'achieved_physical' "
                                 "scales to the positions really sketched. See DESIGN.md and profiles/"}}
    if not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)                        # the CPU arm gets every host core again
        threads = os.cpu_count() or 1
        n_walks = args.cpu_walks or min(threads, args.haps, 16)
        r = run_reference_sample(sg, rd, args, n_walks, args.cpu_coverage, threads)
        line["cpu_baseline"] = {"value": (r["P"] + r["Q"]) / r["t_index"], "unit": UNIT, "cores": threads, "kind": r["kind"],
                                "sample": cpu_sample_desc(n_walks, args.cpu_coverage, args),
                                "read_kmers_per_s": r["Q"] / r["t_reads"] if r["t_reads"] else None,
                                "path_kmers_per_s": r["P"] / r["t_paths"] if r["t_paths"] else None,
                                "index_wall_s_sample": r["t_index"]}
    ix.close()
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_gpu(a)
