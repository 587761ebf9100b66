"""Multi-GPU host logic: sharding of reads / walks over ranks, merge of the per-rank results, and the
torchrun entry point of bench.py for N > 1.  torch.distributed is plumbing only (unique-id broadcast, barrier,
max-over-ranks of the timings); the data exchange itself is NCCL inside libphi_gpu_index.so."""
import ctypes as C
import json
import time

import numpy as np

from . import _abi
from .api import load_library


def split_by_weight(off, world):
    """Contiguous, weight-balanced split of the items described by the offsets array (phi_shard_split_by_weight)."""
    lib = load_library()
    off = np.ascontiguousarray(off, dtype=np.uint64)
    b = np.zeros(world + 1, dtype=np.uint64)
    rc = lib.phi_shard_split_by_weight(off.ctypes.data_as(_abi.u64p), len(off) - 1, world, b.ctypes.data_as(_abi.u64p))
    assert rc == 0
    return b.astype(np.int64)


def owner_of_hash(h, world):
    return load_library().phi_shard_owner_of_hash(int(h), world)


def shard_inputs(graph, reads, rank, world):
    """This rank's shard: contiguous walks (balanced by steps) and contiguous reads (balanced by bases);
    segments and top_order_map are replicated.  Returns (graph_shard, reads_shard, walk_id_base)."""
    wb = split_by_weight(graph.walk_off, world)
    rb = split_by_weight(reads.read_off, world)
    return graph.take_walks(int(wb[rank]), int(wb[rank + 1])), reads.take(int(rb[rank]), int(rb[rank + 1])), int(wb[rank])


def merge_results(parts):
    """Per-rank results -> the global result.  Every rank returns the anchors of ITS walks for all hash ranks, sorted by
    (rank, walk, j); walk ranges ascend with the rank id, so a stable sort of the concatenation on the hash rank alone gives
    the global (rank, walk, j) order (a consumer that fills Anchor_hits[rank][walk] can simply take the parts one after another).
    Per-walk counters are partial sums; n_filtered counts the dropped hash ranks each rank owns."""
    first = parts[0]
    rank = np.concatenate([p.anchor_rank for p in parts])
    walk = np.concatenate([p.anchor_walk for p in parts])
    lens = np.concatenate([np.diff(p.anchor_off.astype(np.int64)) for p in parts])
    vtx = np.concatenate([p.anchor_vtx for p in parts])
    starts = np.concatenate([[0], np.cumsum(lens)])[:-1]
    order = np.argsort(rank, kind="stable")
    from .synth import expand_ranges
    new_vtx = vtx[expand_ranges(starts[order], lens[order])] if len(vtx) else vtx
    return _abi.IndexResultPy(
        count_sp_r=first.count_sp_r, n_walks=first.n_walks, n_filtered=sum(p.n_filtered for p in parts),
        spectrum=first.spectrum,
        anchor_rank=rank[order], anchor_walk=walk[order],
        anchor_off=np.concatenate([[0], np.cumsum(lens[order])]).astype(np.uint64), anchor_vtx=new_vtx.astype(np.int32),
        minimizers_per_walk=np.sum([p.minimizers_per_walk for p in parts], axis=0).astype(np.uint64),
        anchors_per_walk=np.sum([p.anchors_per_walk for p in parts], axis=0).astype(np.uint64),
        read_kmer_positions=sum(p.read_kmer_positions for p in parts), path_kmer_positions=sum(p.path_kmer_positions for p in parts),
        read_minimizers_emitted=sum(p.read_minimizers_emitted for p in parts),
        path_minimizers_emitted=sum(p.path_minimizers_emitted for p in parts), path_hits=sum(p.path_hits for p in parts))


def init_comm(ix, rank, world, walk_id_base, n_walks_global, dist):
    """Create the library's NCCL communicator: rank 0 makes the id, torch.distributed hands it round."""
    box = [ix.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ix.comm_init(rank, world, box[0], walk_id_base, n_walks_global)


# ------------------------------------------------------------------ bench.py entry point for N > 1
def bench_main(args, rank, world, local, B):
    import torch
    import torch.distributed as dist
    import phi_b200
    from . import synth
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    k, w = args.k, args.w
    # weak scaling: every GPU gets a configs[1]-sized shard — its own 49 haplotype walks of the same MHC-shaped graph
    # (49*N haplotypes in total) and its own 10x read set of the same sample (10x*N coverage in total)
    n_haps = args.haps * world
    sg = synth.make_graph(B["SEED"], args.backbone, n_haps, walk_range=(rank * args.haps, (rank + 1) * args.haps))
    rd = synth.make_reads(B["SEED"], sg, args.coverage, read_len=args.read_len, sample_seed=rank)
    g = sg.graph
    ix = phi_b200.PhiGpuIndex(local)
    init_comm(ix, rank, world, rank * args.haps, n_haps, dist)

    def timed(fn):
        t = torch.zeros(1, dtype=torch.float64, device="cuda")
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        t[0] = time.perf_counter() - t0
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                       # max over ranks
        dist.barrier()
        return float(t[0]), out

    ix.upload(g, rd)
    clocks = B["ClockSampler"](local)
    clocks.start()
    for _ in range(args.warmup):
        res = ix.run_resident(k, w, 1.0, download=False)
    stage = []

    def steps():
        r = None
        for _ in range(args.steps):
            r = ix.run_resident(k, w, 1.0, download=False)
            stage.append(ix.times())
        return r
    dt, res = timed(steps)
    gp, rp = ix.pinned_inputs(g, rd)
    for _ in range(args.warmup):
        ix.free_raw(ix.run_raw(gp, rp, k, w, 1.0))

    def e2e_steps():
        n = 0
        for _ in range(args.steps):
            raw = ix.run_raw(gp, rp, k, w, 1.0)
            n = raw.contents.n_anchors
            ix.free_raw(raw)
        return n
    dt_e2e, _ = timed(e2e_steps)
    clk = clocks.stop()
    full = ix.run(gp, rp, k, w, 1.0)
    units = torch.tensor([res.read_kmer_positions + res.path_kmer_positions, res.read_kmer_positions, res.path_kmer_positions,
                          full.n_anchors, full.n_filtered,
                          sum(a.nbytes for a in (g.seg_off, g.seg_bases, g.walk_off, g.walk_vtx, g.top_order_map, rd.read_off, rd.read_bases)),
                          full.wire_bytes()], dtype=torch.float64, device="cuda")
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    tm = {key: float(np.mean([s[key] for s in stage])) for key in stage[0]}
    if rank == 0:
        peak, peak_src = B["measured_peak"]()
        kern = "walk_sketch_kernel" if tm["walk_kernel_ms"] >= tm["read_kernel_ms"] else "read_sketch_kernel"
        alg = B["walk_kernel_algorithmic_bytes"](res, g, k) if kern == "walk_sketch_kernel" else B["read_kernel_algorithmic_bytes"](res, rd)
        achieved = alg / (tm[kern.replace("_sketch_kernel", "_kernel_ms")] * 1e-3) / 1e9
        total = float(units[0])
        line = {"metric": B["METRIC"], "value": total * args.steps / dt, "unit": B["UNIT"], "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8/u64", "data": "synthetic", "config": B["config_dict"](args, world),
                "units_per_step": {"read_kmer_positions": float(units[1]), "path_kmer_positions": float(units[2]),
                                   "spectrum": res.count_sp_r, "anchors": float(units[3]), "filtered_ranks": float(units[4])},
                "stage_ms_rank0": tm,
                "e2e": {"value": total * args.steps / dt_e2e, "unit": B["UNIT"], "h2d_bytes_per_step": int(units[5]),
                        "d2h_bytes_per_step": int(units[6]), "ms_per_step": dt_e2e / args.steps * 1e3,
                        "host_memory": "pinned (phi_gpu_host_alloc) in, pinned (library pool) out"},
                "gpu_launches": int(tm["kernel_launches"]) * args.steps * world, "clocks": clk,
                "exchange": "NCCL: all-to-all of distinct read-minimizer hashes by hash range + broadcast of the sorted slices; "
                            "all-to-all of (rank, count, vertex list) group summaries to the owner of the rank + broadcast of the drop flags; "
                            "anchors stay on the GPU that holds their walk",
                "device_ms_per_step_rank0": tm["total_ms"],
                "roofline": {"bound": "hbm", "kernel": kern + " (rank 0)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": B["kernel_traffic"](kern), "peak_source": peak_src,
                             "kernel_ms": tm[kern.replace("_sketch_kernel", "_kernel_ms")], "sharing": ix.sharing()}}
        line["config"]["workload"] += f"; weak scaling: {args.haps} haplotypes + {args.coverage:g}x reads PER GPU ({n_haps} haplotypes in total)"
        print(json.dumps(line))
    ix.close()
    dist.barrier()
    dist.destroy_process_group()
