#!/usr/bin/env python
"""bench.py — the ILP_index front end on N B200s, one JSON line (driver contract).

A "step" is one pass of the hot path (/root/reference/src/ILP_index.cpp:543-743: read sketch ->
ranked spectrum -> walk sketch + match -> threshold filter -> grouped anchors) over one batch of synthetic
input of a shape BASELINE.json names (--config):

  readme  configs[0]  test/MHC_4.gfa.gz + test/CHM13_reads.fq.gz (the inputs are kept in tests/golden/mhc4.npz)
  c2      configs[1]  MHC-shaped graph, 49 haplotypes x ~5 Mbp, 150 bp reads at --coverage 0.1 / 1 / 10 (default 10)
  c3long  configs[2]  the same graph, 15 kb reads at 10x
  c4      configs[3]  vcf2gfa-shaped graph, 200 haplotypes x 50 Mbp, 10x 150 bp reads            (DEFAULT: the config the
                      metric "... at 1/2/4/8 B200" is quoted on; it fits one GPU)
  c5      configs[4]  500 haplotypes x 150 Mbp, 30x reads, meant for 8 GPUs (--gpus 1 runs the slice rank 0 of 8 would get)

--gpus N > 1 is STRONG scaling: ONE fixed global input, sharded over the ranks (reads by bases, walks by region of the
topological coordinate — phi_b200/multi.py), exchanged inside the library over NCCL.  After the timed region the per-rank
parts are merged (phi_index_result_merge) and a sha256 over the merged result is printed as `result_digest`: it is the
same at every N.  `parity_n` is a bounded case run through the same N-rank path and compared with the CPU oracle.

  value  : (read k-mer positions + path k-mer positions) / s, inputs resident in HBM, K steps bracketed by barrier +
           device synchronise, max over ranks.
  e2e    : the same through phi_gpu_index_run() — host buffers in, host result out, H2D/D2H inside the timed region.
  --impl reference : the UNMODIFIED reference CLI (oracle/_ref/PHI_ref: reference sources + recording Gurobi stub) on
           the box's host cores, front-end stage timed from the reference's own log stamps; readme / c2 / c3long run in
           full, c4 / c5 on a stated slice (the reference's data structures cannot hold them).
"""
import argparse
import hashlib
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

SEED0 = 0x50484931             # "PHI1" + config index
METRIC = "read k-mers counted/s + path k-mers matched/s (ILP_index front end)"
UNIT = "k-mers/s"

CONFIGS = {
    # name: (index in BASELINE.json configs, seed, backbone, haplotypes, read length, coverage, generator keywords, description)
    "readme": dict(idx=0, desc="BASELINE configs[0]: README test, MHC_4.gfa.gz (5 walks x ~5 Mbp) + CHM13_reads.fq.gz (16,401 x 150 bp)"),
    "c2": dict(idx=1, backbone=5_000_000, haps=49, read_len=150, coverage=10.0, gen=dict(founders=8),
               desc="BASELINE configs[1]: synthetic MHC-shaped acyclic graph, 49 haplotypes x ~5 Mbp, nodes chopped to <=30 bp, 150 bp reads"),
    "c3long": dict(idx=2, backbone=5_000_000, haps=49, read_len=15000, coverage=10.0, gen=dict(founders=8), graph_seed_idx=1,
                   desc="BASELINE configs[2]: the configs[1] graph (49 haplotypes x ~5 Mbp) with 15 kb long reads (log-normal, sigma 0.2, 1 % errors)"),
    "c4": dict(idx=3, backbone=50_000_000, haps=200, read_len=150, coverage=10.0, gen=dict(founders=16, sv_frac=0.0, max_indel=50),
               desc="BASELINE configs[3]: vcf2gfa-shaped graph (SNV/indel bubbles, -m 30 chopping), 200 haplotypes x 50 Mbp, 150 bp reads"),
    "c5": dict(idx=4, backbone=150_000_000, haps=500, read_len=150, coverage=30.0, gen=dict(founders=24, sv_frac=0.0, max_indel=50),
               desc="BASELINE configs[4]: chromosome-scale graph, 500 haplotypes x 150 Mbp, 150 bp reads"),
}
BIG = ("c4", "c5")
C5_WORLD = 8                    # configs[4] is defined on 8 GPUs; fewer GPUs run the first ranks' slices of the 8-way partition


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="phi_b200", choices=["phi_b200", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--coverage", type=float, default=None, help="read coverage (default: the config's)")
    ap.add_argument("--haps", type=int, default=None)
    ap.add_argument("--backbone", type=int, default=None)
    ap.add_argument("--founders", type=int, default=None, help="distinct local haplotypes per LD block of the generator")
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("-w", type=int, default=25)
    ap.add_argument("--partition", default="region", choices=["region", "walk"], help="how the walks are sharded over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip digest / parity_n / no-sharing extras (profiling runs)")
    ap.add_argument("--no-digest", action="store_true", help="skip the merged-result digest (the parts of a c5 run are several GB)")
    ap.add_argument("--cpu-walks", type=int, default=0, help="walks in the reference slice of c4/c5 (0: min(nproc, 16))")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    if a.config != "readme":
        a.coverage = c["coverage"] if a.coverage is None else a.coverage
        a.haps = c["haps"] if a.haps is None else a.haps
        a.backbone = c["backbone"] if a.backbone is None else a.backbone
        a.read_len = c["read_len"]
        a.gen = dict(c["gen"])
        if a.founders is not None:
            a.gen["founders"] = a.founders
    return a


def positions(lengths, k, w):
    lengths = np.asarray(lengths, dtype=np.int64)
    return int(np.sum(np.where(lengths >= w + k - 1, lengths - k + 1, 0)))


# ------------------------------------------------------------------ workloads
class Workload:
    """One fixed GLOBAL input + the shard of it a rank works on."""

    def __init__(self, args):
        self.args, self.name = args, args.config
        self.k, self.w, self.T = args.k, args.w, 1.0
        c = CONFIGS[self.name]
        self.seed = SEED0 + c["idx"]
        self.graph_seed = SEED0 + c.get("graph_seed_idx", c["idx"])
        self.sg = None

    def describe(self):
        a, c = self.args, CONFIGS[self.name]
        if self.name == "readme":
            return c["desc"] + f", k={self.k} w={self.w} T={self.T:g}"
        return (c["desc"] + f" at {a.coverage:g}x; generator: {a.haps} haplotypes x {a.backbone / 1e6:g} Mbp backbone, "
                f"{a.gen.get('founders', 8)} distinct local haplotypes per LD block of 400 sites, k={self.k} w={self.w} T={self.T:g}")

    # ---- small configs: the whole input on every rank, then the library's own partition helpers
    def _small(self):
        from phi_b200 import synth
        a = self.args
        if self.name == "readme":
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from golden_cases import Case
            c = Case("mhc4")
            return c.graph, c.reads
        sg = synth.make_graph(self.graph_seed, a.backbone, a.haps, **a.gen)
        if self.name == "c3long":
            rd = synth.make_reads(self.seed, sg, a.coverage, read_len=a.read_len, len_sigma=0.2, sub_err=0.01)
        else:
            rd = synth.make_reads(self.seed, sg, a.coverage, read_len=a.read_len)
        return sg.graph, rd

    def shard(self, rank, world):
        """(graph shard, reads shard, walk_id_base, n_walks_global, region or None, global units (P, Q) or None)."""
        from phi_b200 import multi, synth, _abi
        a = self.args
        if self.name not in BIG:
            g, rd = self._small()
            P = positions(g.walk_lengths(), self.k, self.w)
            Q = positions(np.diff(rd.read_off.astype(np.int64)), self.k, self.w)
            if world == 1:
                return g, rd, 0, g.n_walks, None, (P, Q)
            gs, rs, base, region = multi.shard_inputs(g, rd, rank, world, self.k, self.w, a.partition)
            return gs, rs, base, g.n_walks, region, (P, Q)
        # ---- chromosome-scale configs: never hold the whole walk set; spell walk after walk and keep this rank's part
        part_world, part_rank = world, rank
        if self.name == "c5" and world < C5_WORLD:
            part_world = C5_WORLD                                   # the slice rank `rank` of the 8-way partition would get
        sg = synth.make_graph(self.graph_seed, a.backbone, a.haps, walk_range=(0, 0), **a.gen)
        self.sg = sg
        g0 = sg.graph
        seg_len = np.diff(g0.seg_off.astype(np.int64))
        seq = synth.mosaic_sequence(self.seed, sg)
        n_reads = synth.n_reads_big(len(seq), a.coverage, a.read_len)
        rlo, rhi = n_reads * part_rank // part_world, n_reads * (part_rank + 1) // part_world
        rd = synth.make_reads_big(self.seed, sg, a.coverage, read_len=a.read_len, read_range=(rlo, rhi), seq=seq)
        Q = n_reads * (min(a.read_len, len(seq)) - self.k + 1)
        del seq
        region, base = None, 0
        bounds = None
        if part_world > 1 and a.partition == "region":
            sample = [sg.walk_of(sg.allele_row(h)) for h in range(0, a.haps, max(1, a.haps // 4))][:4]
            gsample = _abi.Graph(g0.seg_off, g0.seg_bases, np.concatenate([[0], np.cumsum([len(x) for x in sample])]),
                                 np.concatenate(sample), g0.top_order_map)
            bounds = multi.region_bounds(gsample, part_world)
            region = (int(bounds[part_rank]), int(bounds[part_rank + 1]))
        walks, P = [], 0
        wlo, whi = 0, a.haps
        if part_world > 1 and a.partition == "walk":
            wlo, whi = a.haps * part_rank // part_world, a.haps * (part_rank + 1) // part_world
            base = wlo
        s_lo, s_hi = 0, sg.n_sites
        if region is not None:
            # the generator numbers the vertices along the backbone and top_order_map is the identity: the topological coordinate of a
            # vertex is its offset in the segment store, so the region is a vertex range, i.e. a range of variant sites (+ a margin that
            # holds far more than the w / k-1 bases of context the slicer keeps)
            v_lo = int(np.searchsorted(g0.seg_off, np.uint64(region[0]), side="left"))
            v_hi = int(np.searchsorted(g0.seg_off, np.uint64(min(region[1], int(g0.seg_off[-1]))), side="left"))
            pf = sg.piece_first_node
            s_lo = max(0, int(np.searchsorted(pf, v_lo, side="right") - 1) // 3 - 64)
            s_hi = min(sg.n_sites, int(np.searchsorted(pf, v_hi, side="right")) // 3 + 64)
        BATCH = 8
        for h0 in range(0, a.haps, BATCH):
            hs = range(h0, min(h0 + BATCH, a.haps))
            if region is not None:
                full = [sg.walk_of_hap_sites(h, s_lo, s_hi) for h in hs]
                for h in hs:
                    P += positions([sg.hap_length(h)], self.k, self.w)
            else:
                full = [sg.walk_of(sg.allele_row(h)) for h in hs]
                for x in full:
                    P += positions([int(seg_len[x].sum())], self.k, self.w)
            if region is not None:
                gb = _abi.Graph(g0.seg_off, g0.seg_bases, np.concatenate([[0], np.cumsum([len(x) for x in full])]),
                                np.concatenate(full), g0.top_order_map)
                sl = multi.slice_walks(gb, self.k, self.w, region[0], region[1])
                so = sl.walk_off.astype(np.int64)
                walks.extend(sl.walk_vtx[so[i]:so[i + 1]] for i in range(len(full)))
            else:
                walks.extend(x for i, x in enumerate(full) if wlo <= h0 + i < whi)
        gs = _abi.Graph(g0.seg_off, g0.seg_bases, np.concatenate([[0], np.cumsum([len(x) for x in walks])]).astype(np.uint64),
                        np.concatenate(walks).astype(np.uint32) if walks else np.zeros(0, dtype=np.uint32), g0.top_order_map)
        units = (P, Q)
        if self.name == "c5" and world < C5_WORLD:
            units = None                                            # a slice: the units are what the ranks report
        return gs, rd, base, a.haps, region, units

    # ---- what the reference arm (and the cpu_baseline leg) runs
    def reference_input(self, threads):
        """(graph, reads, sample description, full?) for the CPU reference."""
        from phi_b200 import synth
        a = self.args
        if self.name not in BIG:
            g, rd = self._small()
            return g, rd, "the full configuration (all walks, all reads); front-end stage only (log stamps 'Graph has' -> 'Filtered/Retained')", True
        n_walks = a.cpu_walks or min(threads, a.haps, 16)
        frac = 10 if self.name == "c4" else 30
        sg = synth.make_graph(self.graph_seed, a.backbone // frac, a.haps, walk_range=(0, n_walks), **a.gen)
        rd = synth.make_reads_big(self.seed, sg, a.coverage if self.name == "c4" else a.coverage / 3, read_len=a.read_len)
        desc = (f"SLICE (the reference's kmer_index / Anchor_hits / in_paths cannot hold this config, BASELINE.md §3): the same generator on 1/{frac} of the "
                f"backbone ({a.backbone // frac / 1e6:g} Mbp), {n_walks} of its {a.haps} walks, {rd.n_reads} reads of {a.read_len} bp; per-k-mer rate, "
                f"EXTRAPOLATED to the full config; front-end stage only (log stamps 'Graph has' -> 'Filtered/Retained')")
        return sg.graph, rd, desc, False


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


# ------------------------------------------------------------------ CPU reference
def run_reference(g, rd, k, w, threads, files=None):
    """Times the reference's own front end (log-stamp deltas) on the given graph and reads.  Kills the process once
    the front end is done (model construction is not timed).  files: (gfa, fasta) already written for this input."""
    from phi_b200 import synth
    exe = os.path.join(ROOT, "oracle", "_ref", "PHI_ref")
    P = positions(g.walk_lengths(), k, w)
    Q = positions(np.diff(rd.read_off.astype(np.int64)), k, w)
    if not os.path.exists(exe):
        # the reference did not compile here: time the plain-C port instead (kind "port")
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import phi_io
        t0 = time.time()
        phi_io.oracle_index(g, rd, k, w, 1.0, threads)
        dt = time.time() - t0
        return dict(kind="port", P=P, Q=Q, t_index=dt, t_reads=None, t_paths=None)
    with tempfile.TemporaryDirectory() as tmp:
        if files is None:
            gfa, fa = os.path.join(tmp, "g.gfa"), os.path.join(tmp, "r.fa")
            synth.write_gfa(g, gfa)
            synth.write_fasta(rd, fa)
        else:
            gfa, fa = files
        env = dict(os.environ, PHI_STUB_DUMP=os.path.join(tmp, "dump.txt"))
        p = subprocess.Popen([exe, "-g", gfa, "-r", fa, "-o", os.path.join(tmp, "o.fa"), "-t", str(threads), "-k", str(k),
                              "-w", str(w)], env=env, stderr=subprocess.PIPE, stdout=subprocess.DEVNULL, text=True)
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


def config_dict(wl, n):
    return {"workload": wl.describe(), "seed": wl.seed, "gpus": n,
            "l2": "no explicit flush: the per-step working set (walk steps, step offsets, chunk table, reads, spectrum table, hit and "
                  "group buffers: > 400 MB per GPU) exceeds the 126 MB L2"}


def main_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = Workload(args)
    threads = os.cpu_count() or 1
    g, rd, desc, full = wl.reference_input(threads)
    vals, times, r = [], [], None
    from phi_b200 import synth
    with tempfile.TemporaryDirectory() as tmp:                    # the input files are written once, every step is one run of the reference CLI
        files = (os.path.join(tmp, "g.gfa"), os.path.join(tmp, "r.fa"))
        synth.write_gfa(g, files[0])
        synth.write_fasta(rd, files[1])
        for i in range(args.warmup + args.steps):
            r = run_reference(g, rd, args.k, args.w, threads, files)
            if i >= args.warmup:
                vals.append((r["P"] + r["Q"]) / r["t_index"])
                times.append(r["t_index"])
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(times)) * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8/u64", "data": "synthetic" if args.config != "readme" else "README test files",
            "config": config_dict(wl, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": r["kind"], "sample": desc, "full_config": full,
                             "units_per_step": {"read_kmer_positions": r["Q"], "path_kmer_positions": r["P"]},
                             "read_kmers_per_s": r["Q"] / r["t_reads"] if r["t_reads"] else None,
                             "path_kmers_per_s": r["P"] / r["t_paths"] if r["t_paths"] else None},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


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


def kernel_profile(name, config):
    """ncu numbers of the current build for this config (profiles/kernel_traffic.json: written by profiles/make_kernel_traffic.py from
    an `ncu --set full` capture of the same bench command; re-captured per round, version inside)."""
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    try:
        j = json.load(open(tp)).get(config, {})
        return j.get(name, {}), j.get("source")
    except Exception:
        return {}, None


def result_digest(res):
    """sha256 over the merged result in global order: the ranked spectrum, first group of every rank, the vertex lists, the member
    walks, per-walk counters, filtered ranks.  Independent of the number of GPUs."""
    h = hashlib.sha256()
    for a, dt in ((res.spectrum, np.uint64), (res.rank_off if res.rank_off is not None else np.zeros(1), np.uint32), (res.group_len, np.uint8),
                  (res.group_vtx, np.int32), (res.group_member_off, np.uint32), (res.member_walk, np.int32),
                  (res.minimizers_per_walk, np.uint64), (res.anchors_per_walk, np.uint64)):
        h.update(np.ascontiguousarray(a, dtype=dt).tobytes())
    h.update(np.array([res.count_sp_r, res.n_filtered, res.n_walks], dtype=np.int64).tobytes())
    return h.hexdigest()


class Dist:
    """torch.distributed plumbing (barrier, max / sum over ranks, object hand-round); a no-op at world 1."""

    def __init__(self, rank, world, local):
        self.rank, self.world = rank, world
        if world > 1:
            # (the library sets these before it creates its communicator; here torch's process group initialises NCCL first and NCCL
            # reads its environment once per process)
            os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "8")
            os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "32")
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(local)
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
            self.torch, self.dist = torch, dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def reduce(self, values, op="sum"):
        if self.world == 1:
            return [float(v) for v in values]
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def bcast_obj(self, obj):
        if self.world == 1:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def gather_results(self, part, tag):
        """Per-rank results -> list on rank 0 (through /dev/shm: the parts can be hundreds of MB)."""
        if self.world == 1:
            return [part]
        import pickle
        import shutil
        need = 2 * self.world * (part.wire_bytes() + (1 << 20))
        base = next((b for b in ("/dev/shm", tempfile.gettempdir()) if os.path.isdir(b) and shutil.disk_usage(b).free > need), tempfile.gettempdir())
        d = os.path.join(base, f"phi_bench_{os.environ.get('MASTER_PORT', '0')}_{tag}")
        ok = 1.0
        try:
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, f"part{self.rank}.pkl"), "wb") as f:
                pickle.dump(part, f, protocol=4)
        except Exception:
            ok = 0.0
        ok = min(self.reduce([ok], "sum")[0] / self.world, ok)       # every rank takes the same path (the reduce is the barrier)
        parts = None
        if self.rank == 0 and ok == 1.0:
            try:
                parts = []
                for r in range(self.world):
                    with open(os.path.join(d, f"part{r}.pkl"), "rb") as f:
                        parts.append(pickle.load(f))
            except Exception:
                parts = None
        self.dist.barrier()
        shutil.rmtree(d, ignore_errors=True) if self.rank == 0 else None
        return parts

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def setup_index(local, rank, world, base, n_walks_global, region, D):
    import phi_b200
    ix = phi_b200.PhiGpuIndex(local)
    if world > 1:
        uid = D.bcast_obj(ix.comm_unique_id() if rank == 0 else None)
        ix.comm_init(rank, world, uid, base, n_walks_global)
    if region is not None:
        ix.set_walk_region(region[0], region[1])
    return ix


def parity_case(rank, world, local, D, partition):
    """A bounded case (dirty bytes included) through the same N-rank path, merged and compared with the CPU oracle on the unsharded
    input.  The oracle is the checker here, nothing else (oracle/phi_oracle.h)."""
    from phi_b200 import synth, multi
    sg = synth.make_graph(4242, 400_000, 11, lower_frac=0.004, n_frac=0.001)
    rd = synth.make_reads(4242, sg, 3.0, lower_frac=0.004, n_frac=0.001)
    k, w, T = 31, 25, 0.9
    if world == 1:
        gs, rs, base, region = sg.graph, rd, 0, None
    else:
        gs, rs, base, region = multi.shard_inputs(sg.graph, rd, rank, world, k, w, partition)
    ix = setup_index(local, rank, world, base, sg.graph.n_walks, region, D)
    part = ix.run(gs, rs, k, w, T, expand=False)
    ix.close()
    parts = D.gather_results(part, "parity")
    verdict = None
    if rank == 0 and parts is None:
        verdict = "not checked (no scratch space for the parts)"
    elif rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import phi_io
        got = multi.merge_results(parts)
        want = phi_io.oracle_index(sg.graph, rd, k, w, T)
        ok = (got.count_sp_r == want.count_sp_r and got.n_filtered == want.n_filtered
              and all(np.array_equal(getattr(got, f), getattr(want, f)) for f in
                      ("spectrum", "anchor_rank", "anchor_walk", "anchor_off", "anchor_vtx", "minimizers_per_walk", "anchors_per_walk"))
              and got.path_kmer_positions == want.path_kmer_positions and got.read_kmer_positions == want.read_kmer_positions)
        verdict = "ok" if ok else "MISMATCH"
    return D.bcast_obj(verdict)


def main_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    all_cpus = os.sched_getaffinity(0)
    affinity = bind_to_gpu_numa(local)
    D = Dist(rank, world, local)
    wl = Workload(args)
    k, w, T = wl.k, wl.w, wl.T
    t_gen = time.time()
    g, rd, base, n_walks_global, region, units_global = wl.shard(rank, world)
    t_gen = time.time() - t_gen
    ix = setup_index(local, rank, world, base, n_walks_global, region, D)

    def timed(fn):
        D.barrier()
        t0 = time.perf_counter()
        out = fn()                                           # every run ends with a stream synchronise inside the library
        dt = time.perf_counter() - t0
        dt = D.reduce([dt], "max")[0]                        # max over ranks
        D.barrier()
        return dt, out

    # ---- resident: inputs already in HBM when the timed region starts
    ix.upload(g, rd)
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(args.warmup):
        res = ix.run_resident(k, w, T, download=False)
    stage = []

    def steps():
        r = None
        for _ in range(args.steps):
            r = ix.run_resident(k, w, T, download=False)
            stage.append(ix.times())
        return r
    dt, res = timed(steps)
    share = ix.sharing()
    tm = {key: float(np.mean([s[key] for s in stage])) for key in stage[0]}

    # ---- end to end through the drop-in call: pinned host buffers -> host result; every step copies all inputs host->device
    # and the whole result device->host inside the timed region
    gp, rp = ix.pinned_inputs(g, rd)
    for _ in range(args.warmup):
        ix.free_raw(ix.run_raw(gp, rp, k, w, T))
    e2e_stage = []

    def e2e_steps():
        n = 0
        for _ in range(args.steps):
            raw = ix.run_raw(gp, rp, k, w, T)
            n = raw.contents.n_anchors                       # the caller reads the result
            ix.free_raw(raw)
            e2e_stage.append(ix.times())
        return n
    dt_e2e, n_anchors_seen = timed(e2e_steps)
    clk = clocks.stop()
    full = ix.run(gp, rp, k, w, T, expand=False)
    assert full.n_anchors == n_anchors_seen
    h2d = sum(a.nbytes for a in (g.seg_off, g.seg_bases, g.walk_off, g.walk_vtx, g.top_order_map, rd.read_off, rd.read_bases))
    d2h = full.wire_bytes()

    # ---- sums over the ranks
    loc = [res.read_kmer_positions, res.path_kmer_positions, res.read_minimizers_emitted, full.path_minimizers_emitted, res.path_hits,
           h2d, d2h, share["unique_windows"], share["unique_hits"], len(g.walk_vtx), int(tm["kernel_launches"]), share["chunks"], share["tiles"]]
    tot = D.reduce(loc, "sum")
    Q_run, P_run = tot[0], tot[1]
    if units_global is not None:
        assert (int(P_run), int(Q_run)) == (units_global[0], units_global[1]), ("units", P_run, Q_run, units_global)
    units = P_run + Q_run
    tmax = dict(zip(tm, D.reduce(list(tm.values()), "max")))

    extras = {}
    if not args.no_extras:
        # ---- the merged result and its digest (the same at every N), the N-rank parity case
        t0 = time.time()
        parts = D.gather_results(full, "digest") if not args.no_digest else None
        if rank == 0 and parts is None and not args.no_digest:
            extras["result_digest"] = None                   # (the parts could not be handed to rank 0: no scratch space)
        if rank == 0 and parts is not None:
            from phi_b200 import multi
            t1 = time.time()
            merged = multi.merge_results(parts, expand=False) if world > 1 else full
            extras["merge_ms"] = getattr(multi.merge_results, "last_call_ms", 0.0) if world > 1 else 0.0   # phi_index_result_merge alone
            extras["merge_with_numpy_conversions_ms"] = (time.time() - t1) * 1e3
            extras["result_digest"] = result_digest(merged)
            extras["result"] = {"spectrum": merged.count_sp_r, "groups": merged.n_groups, "anchors": int(len(merged.member_walk)),
                                "group_vertices": int(len(merged.group_vtx)), "filtered_ranks": merged.n_filtered}
            del merged
        del parts
        extras["parity_n"] = parity_case(rank, world, local, D, args.partition)
        # ---- the same resident step with walk sharing switched off (every chunk of every walk sketched on its own), when it fits
        est_hits = tot[1] / max(1, world) * 2.0 / (w + 1.0) * 1.3
        if est_hits * 40 < 60e9 and est_hits < 3.5e9:
            ix.set_walk_sharing(share=False)
            for _ in range(2):
                ix.run_resident(k, w, T, download=False)
            st2 = []

            def ns_steps():
                for _ in range(3):
                    ix.run_resident(k, w, T, download=False)
                    st2.append(ix.times())
            dt_ns, _ = timed(ns_steps)
            extras["no_sharing"] = {"value": units * 3 / dt_ns, "ms_per_step": dt_ns / 3 * 1e3,
                                    "walk_kernel_ms": float(np.mean([s["walk_kernel_ms"] for s in st2])),
                                    "note": "walk sharing off (phi_gpu_index_set_walk_sharing share=0): every chunk of every walk is sketched; same result"}
            ix.set_walk_sharing(share=True)
        else:
            extras["no_sharing"] = None

    peak, peak_src = measured_peak()
    line = None
    if rank == 0:
        # ---- roofline of the dominant sketch kernel (rank 0's launch), on the units THE LAUNCH physically processes
        steps_r0, P_r0 = len(g.walk_vtx), max(1, res.path_kmer_positions)
        uf = share["unique_windows"] / P_r0
        mean_node = max(1.0, float(np.diff(g.seg_off.astype(np.int64))[g.walk_vtx[:: max(1, len(g.walk_vtx) // 1_000_000)]].mean())) if steps_r0 else 1.0
        hit_vtx = share["unique_hits"] * (1.0 + (k - 1) / mean_node)
        walk_alg = (0.25 * share["unique_windows"] + 4.0 * steps_r0 * uf + 32.0 * full.path_minimizers_emitted * uf
                    + 16.0 * share["unique_hits"] + 4.0 * hit_vtx)
        walk_eff = walk_alg / max(uf, 1e-12) if uf > 0 else 0.0
        read_alg = float(rd.read_bases.nbytes) + 64.0 * res.read_minimizers_emitted
        kernels = {}
        for name, ms, alg, nunits in (("walk_sketch_kernel", tm["walk_kernel_ms"], walk_alg, share["unique_windows"]),
                                      ("read_sketch_kernel", tm["read_kernel_ms"], read_alg, res.read_kmer_positions)):
            prof, _ = kernel_profile(name, args.config)
            ach = alg / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            kernels[name] = {"kernel_ms": ms, "units_per_launch": int(nunits), "algorithmic_bytes_per_launch": alg, "achieved": ach, "frac": ach / peak,
                             "traffic": prof.get("dram_bytes_per_launch"), "issue_slot_util": prof.get("issue_slot_util"),
                             "warp_inst_per_kmer": prof.get("warp_inst_per_kmer")}
        kernels["walk_sketch_kernel"]["effective"] = {
            "achieved": walk_eff / (tm["walk_kernel_ms"] * 1e-3) / 1e9 if tm["walk_kernel_ms"] > 0 else 0.0,
            "note": "bytes of ALL path k-mer positions the launch accounts for (identical chunks are sketched once): not a roofline figure"}
        dom = max(kernels, key=lambda n: kernels[n]["kernel_ms"])
        _, prof_src = kernel_profile(dom, args.config)
        line = {"metric": METRIC, "value": units * args.steps / dt, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u8/u64", "data": "synthetic" if args.config != "readme" else "README test files", "config": config_dict(wl, world),
                "read_kmers_per_s": Q_run / ((tmax["read_sketch_ms"] + tmax["spectrum_ms"]) * 1e-3),
                "path_kmers_per_s": P_run / ((tmax["graph_prep_ms"] + tmax["walk_sketch_ms"] + tmax["filter_ms"]) * 1e-3),
                "index_wall_s": dt_e2e / args.steps,
                "units_per_step": {"read_kmer_positions": Q_run, "path_kmer_positions": P_run, "read_minimizers": tot[2], "path_minimizers": tot[3],
                                   "path_hits": tot[4], "spectrum": res.count_sp_r, "walk_steps_uploaded": tot[9]},
                "stage_ms_rank0": tm, "stage_ms_max": tmax,
                "e2e": {"value": units * args.steps / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": int(tot[5]), "d2h_bytes_per_step": int(tot[6]),
                        "ms_per_step": dt_e2e / args.steps * 1e3, "host_memory": "pinned (phi_gpu_host_alloc) in, pinned (library pool) out",
                        "stage_ms_rank0": {key: float(np.mean([x[key] for x in e2e_stage])) for key in e2e_stage[0]},
                        "note": "per-rank parts; the host-side merge of the parts (merge_ms, N > 1 only) is outside the timed region"},
                "gpu_launches": int(tot[10]) * args.steps,
                "clocks": clk, "cpu_affinity": affinity, "partition": (args.partition if world > 1 else "none"), "generate_s": t_gen,
                "device_ms_per_step_rank0": tm["total_ms"],
                "sharing": {"chunks": tot[11], "tiles": tot[12], "unique_windows": tot[7], "unique_hits": tot[8], "path_kmer_positions": P_run,
                            "unique_fraction": tot[7] / max(1.0, P_run)},
                "roofline": {"bound": "hbm", "kernel": dom + (" (rank 0)" if world > 1 else ""), "achieved": kernels[dom]["achieved"], "peak": peak,
                             "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": kernels[dom]["traffic"], "peak_source": peak_src,
                             "traffic_source": prof_src,
                             "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes_per_launch"], "kernel_ms": kernels[dom]["kernel_ms"],
                             "kernels": kernels,
                             "note": "achieved = SURVEY §8(d) bytes of the units the launch physically processes (walk kernel: the windows of the representative "
                                     "chunks only) / the kernel's CUDA-event time. Both sketch kernels are integer-issue bound (issue_slot_util, "
                                     "warp_inst_per_kmer from ncu), so their HBM fraction is low by construction. See DESIGN.md §3.4 and profiles/"}}
        line.update(extras)
        if world > 1:
            line["exchange"] = ("NCCL inside the library: all-to-all of the locally distinct read-minimizer hashes by hash range + all-gather of the owners' sorted "
                                "slices; all-to-all of one (rank, count, vertex list) summary per local group to the owner of the rank + all-reduce of the drop flags")
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        os.sched_setaffinity(0, all_cpus)                    # the CPU arm gets every host core again
        threads = os.cpu_count() or 1
        cg, crd, desc, fullcfg = wl.reference_input(threads)
        r = run_reference(cg, crd, k, w, threads)
        line["cpu_baseline"] = {"value": (r["P"] + r["Q"]) / r["t_index"], "unit": UNIT, "cores": threads, "kind": r["kind"], "sample": desc,
                                "full_config": fullcfg, "units": {"read_kmer_positions": r["Q"], "path_kmer_positions": r["P"]},
                                "read_kmers_per_s": r["Q"] / r["t_reads"] if r["t_reads"] else None,
                                "path_kmers_per_s": r["P"] / r["t_paths"] if r["t_paths"] else None,
                                "index_wall_s_sample": r["t_index"]}
    ix.close()
    if rank == 0:
        print(json.dumps(line))
    D.close()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_gpu(a)
