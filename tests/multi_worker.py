"""Worker launched by the multi-rank tests:  python -m torch.distributed.run ... tests/multi_worker.py {gpu|cpu} <case>

gpu : every rank drives one B200 through the C ABI (NCCL exchange inside the library), rank 0 merges and checks
      against the oracle on the unsharded input.
cpu : gloo, no GPU — the same partition functions (phi_shard_*: whole walks, or every walk cut to a region of the graph) and
      the library's merge (phi_index_result_merge), with the per-rank compute done by the oracle and the exchanges done with
      torch.distributed object collectives: a check of the sharded ALGORITHM (hash-range ownership, rank offsets, which rank
      owns which window of a sliced walk, group summaries to the owner, owner-side threshold, shared drop flags, local
      ordering, merge), not of the kernels.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch.distributed as dist  # noqa: E402

import phi_io  # noqa: E402
from golden_cases import Case, assert_same_result  # noqa: E402
from phi_b200 import _abi, multi  # noqa: E402


def load(name):
    if name.startswith("synth:"):
        from phi_b200 import synth
        _, seed, backbone, haps, cov = name.split(":")
        sg = synth.make_graph(int(seed), int(backbone), int(haps))
        rd = synth.make_reads(int(seed), sg, float(cov))
        return sg.graph, rd, 31, 25, 1.0
    c = Case(name)
    return c.graph, c.reads, c.k, c.w, c.T


def main_gpu(name, mode="region"):
    import torch
    import phi_b200
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    debug = 1 if name.endswith("+d1") else 0             # "+d1": with the shared k-mer statistic (ILP_index.cpp:565-606)
    name = name[:-3] if debug else name
    g, rd, k, w, T = load(name)
    gs, rs, base, region = multi.shard_inputs(g, rd, rank, world, k, w, mode)
    ix = phi_b200.PhiGpuIndex(local)
    multi.init_comm(ix, rank, world, base, g.n_walks, dist)
    if region is not None:
        ix.set_walk_region(*region)
    part = ix.run(gs, rs, k, w, T, debug=debug)
    part2 = ix.run(gs, rs, k, w, T, debug=debug)         # repeatable
    assert_same_result(part, part2)
    assert rank == 0 or len(part.spectrum) == 0          # the ranked spectrum is copied out on rank 0 only
    parts = [None] * world
    dist.gather_object(part, parts if rank == 0 else None, dst=0)
    if rank == 0:
        got = multi.merge_results(parts)
        want = phi_io.oracle_index(g, rd, k, w, T, debug=debug)
        got.path_hits = want.path_hits                   # pre-filter hit count is per-GPU bookkeeping
        assert_same_result(want, got)
        if debug:
            assert parts[0].n_walk_kmers == want.n_walk_kmers and all(np.array_equal(p.shared_kmer_hist, want.shared_kmer_hist) for p in parts)
        print(f"MULTI_OK gpu world={world} case={name} mode={mode} region={region} spectrum={got.count_sp_r} anchors={got.n_anchors}")
    ix.close()
    dist.barrier()


def py_filter(hits, n_walks_global, T):
    """hits: list of (rank, walk, seq, vertices).  Reference filter + order (ILP_index.cpp:670-716) for the ranks present."""
    by_rank = {}
    for h in hits:
        by_rank.setdefault(h[0], []).append(h)
    out, filtered = [], 0
    thr = np.float32(T) * np.float32(n_walks_global)
    for r in sorted(by_rank):
        hs = sorted(by_rank[r], key=lambda x: (x[1], x[2]))            # insertion order: walk asc, path order
        groups = {}
        for h in hs:
            groups.setdefault("".join(f"{v}_" for v in h[3]), []).append(h)
        if any(np.float32(len(m)) >= thr for m in groups.values()):
            filtered += 1
            continue
        per_walk = {}
        for key in sorted(groups):                                      # std::map<std::string> order
            for h in groups[key]:
                per_walk.setdefault(h[1], []).append(h)
        for wk in sorted(per_walk):
            out.extend(per_walk[wk])
    return out, filtered


def main_cpu(name, mode="walk"):
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    g, rd, k, w, T = load(name)
    gs, rs, base, region = multi.shard_inputs(g, rd, rank, world, k, w, mode)
    empty = _abi.Graph(g.seg_off, g.seg_bases, np.zeros(1, dtype=np.uint64), np.zeros(0, dtype=np.uint32), g.top_order_map)
    local = phi_io.oracle_index(empty, rs, k, w, T).spectrum            # this rank's distinct read-minimizer hashes
    owner = np.array([multi.owner_of_hash(h, world) for h in local], dtype=np.int64)
    assert np.all(np.diff(owner) >= 0)                                  # range partition: owners ascend with the hash
    send = [local[owner == o] for o in range(world)]
    allsend = [None] * world
    dist.all_gather_object(allsend, send)
    mine = np.unique(np.concatenate([allsend[src][rank] for src in range(world)]))
    slices = [None] * world
    dist.all_gather_object(slices, mine)
    spectrum = np.concatenate(slices)                                   # concatenation of range slices is globally sorted
    own_off = np.concatenate([[0], np.cumsum([len(s) for s in slices])])
    assert np.all(spectrum[1:] > spectrum[:-1])
    if region is None:
        sk, hashes = phi_io.oracle_sketch_walks(gs, k, w)
        mpw_local = sk.minimizers_per_walk
    else:
        # region partition: this rank holds every walk cut to its region plus context; of what the slices emit it owns the minimizers
        # whose emitting window ends (last k-mer starts) on a vertex inside the region — exactly what the library's chunks own
        sk, hashes, owner = phi_io.oracle_sketch_walks_owner(gs, k, w)
        seg_len = np.diff(g.seg_off.astype(np.int64))
        order = np.argsort(g.top_order_map, kind="stable")
        coord = np.zeros(g.n_vtx, dtype=np.float64)
        coord[order] = np.concatenate([[0], np.cumsum(seg_len[order])])[:-1]
        own = (coord[owner] >= float(region[0])) & (coord[owner] < float(region[1]))
        mpw_local = np.bincount(sk.anchor_walk[own], minlength=gs.n_walks).astype(np.uint64)
    idx = np.searchsorted(spectrum, hashes)
    idx[idx == len(spectrum)] = 0
    hit = spectrum[idx] == hashes if len(spectrum) else np.zeros(len(hashes), dtype=bool)
    if region is not None:
        hit &= own
    off = sk.anchor_off.astype(np.int64)
    hits = [(int(idx[a]), int(sk.anchor_walk[a]) + base, int(a), sk.anchor_vtx[off[a]:off[a + 1]].tolist()) for a in np.nonzero(hit)[0]]
    # group summaries (rank, vertex-list key) -> count travel to the owner of the rank; the owner adds them up and decides
    groups = {}
    for h in hits:
        gk = (h[0], "".join(f"{v}_" for v in h[3]))
        groups[gk] = groups.get(gk, 0) + 1
    route = [{gk: c for gk, c in groups.items() if own_off[o] <= gk[0] < own_off[o + 1]} for o in range(world)]
    allroute = [None] * world
    dist.all_gather_object(allroute, route)
    total = {}
    for src in range(world):
        for gk, c in allroute[src][rank].items():
            total[gk] = total.get(gk, 0) + c
    thr = np.float32(T) * np.float32(g.n_walks)
    dropped_here = sorted({gk[0] for gk, c in total.items() if np.float32(c) >= thr})
    alldrop = [None] * world
    dist.all_gather_object(alldrop, dropped_here)                       # the drop flags are shared
    dropped = set(r for d in alldrop for r in d)
    filtered = len(dropped_here)
    kept, none_dropped = py_filter([h for h in hits if h[0] not in dropped], 1 << 30, 1.0)   # local order of the local walks' anchors
    assert none_dropped == 0
    mpw = np.zeros(g.n_walks, dtype=np.uint64)
    mpw[base:base + gs.n_walks] = mpw_local
    apw = np.bincount([h[1] for h in kept], minlength=g.n_walks).astype(np.uint64)
    voff = np.concatenate([[0], np.cumsum([len(h[3]) for h in kept])]).astype(np.uint64)
    part = _abi.IndexResultPy(
        count_sp_r=len(spectrum), n_walks=g.n_walks, n_filtered=filtered, spectrum=spectrum,
        anchor_rank=np.array([h[0] for h in kept], dtype=np.int32), anchor_walk=np.array([h[1] for h in kept], dtype=np.int32),
        anchor_off=voff, anchor_vtx=np.array([v for h in kept for v in h[3]], dtype=np.int32),
        minimizers_per_walk=mpw, anchors_per_walk=apw)
    # the ABI's grouped form of this rank's anchors (what the library returns and phi_index_result_merge takes)
    import dataclasses
    ro, gl, gv, mo, mw = phi_io.group_anchors(part)
    part = dataclasses.replace(part, rank_off=ro, group_len=gl, group_vtx=gv, group_member_off=mo, member_walk=mw, n_groups=len(gl),
                               spectrum=spectrum if rank == 0 else np.zeros(0, dtype=np.uint64))
    parts = [None] * world
    dist.gather_object(part, parts if rank == 0 else None, dst=0)
    if rank == 0:
        got = multi.merge_results(parts)
        want = phi_io.oracle_index(g, rd, k, w, T)
        for f in ("spectrum", "anchor_rank", "anchor_walk", "anchor_off", "anchor_vtx", "minimizers_per_walk", "anchors_per_walk"):
            assert np.array_equal(getattr(got, f), getattr(want, f)), f
        assert got.n_filtered == want.n_filtered and got.count_sp_r == want.count_sp_r
        print(f"MULTI_OK cpu world={world} case={name} mode={mode} spectrum={got.count_sp_r} anchors={got.n_anchors}")
    dist.barrier()


if __name__ == "__main__":
    if sys.argv[1] == "gpu":
        main_gpu(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "region")
    else:
        main_cpu(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "walk")
