"""CPU: the host side of the multi-GPU path — region partition of the walks (phi_shard_walk_regions / phi_shard_slice_walks) and the
merge of per-GPU parts (phi_index_result_merge), checked against the oracle's result re-partitioned on the host."""
import dataclasses

import numpy as np
import pytest

import phi_io
from phi_b200 import _abi, multi, synth


def grouped(res):
    ro, gl, gv, mo, mw = phi_io.group_anchors(res)
    return dataclasses.replace(res, rank_off=ro, group_len=gl, group_vtx=gv, group_member_off=mo, member_walk=mw, n_groups=len(gl))


def take_anchors(res, idx, **over):
    lens = np.diff(res.anchor_off.astype(np.int64))
    starts = res.anchor_off.astype(np.int64)[idx]
    vtx = np.concatenate([res.anchor_vtx[s:s + n] for s, n in zip(starts, lens[idx])]) if len(idx) else np.zeros(0, dtype=np.int32)
    return dataclasses.replace(res, anchor_rank=res.anchor_rank[idx], anchor_walk=res.anchor_walk[idx],
                               anchor_off=np.concatenate([[0], np.cumsum(lens[idx])]).astype(np.uint64), anchor_vtx=vtx.astype(np.int32),
                               anchors_per_walk=np.bincount(res.anchor_walk[idx], minlength=res.n_walks).astype(np.uint64), **over)


@pytest.fixture(scope="module")
def case():
    sg = synth.make_graph(42, 60000, 9, founders=4, block_sites=30)
    rd = synth.make_reads(42, sg, 6.0)
    return sg.graph, rd, phi_io.oracle_index(sg.graph, rd, 31, 25, 1.0)


@pytest.mark.parametrize("n_parts,seed,threads", [(1, 0, 1), (2, 1, 1), (3, 2, 1), (8, 3, 1), (1, 4, 3), (3, 5, 5), (8, 6, 16)])
def test_merge_of_arbitrary_parts_is_the_whole(case, n_parts, seed, threads, monkeypatch):
    """Anchors dealt to the parts at random: the same (rank, vertex list) group then lives in several parts with different member
    walks, the groups of one rank are spread over the parts — the merge must restore key order and unite the members.  Big merges
    cut the hash ranks into blocks and run them on several threads (count pass, prefix, write pass): PHI_MERGE_THREADS forces that
    path on this small case too."""
    g, rd, want = case
    monkeypatch.setenv("PHI_MERGE_THREADS", str(threads))
    rng = np.random.default_rng(seed)
    assign = rng.integers(0, n_parts, want.n_anchors)
    parts = []
    for p in range(n_parts):
        parts.append(grouped(take_anchors(
            want, np.nonzero(assign == p)[0], n_filtered=want.n_filtered if p == n_parts - 1 else 0,
            minimizers_per_walk=want.minimizers_per_walk if p == 0 else want.minimizers_per_walk * 0,
            spectrum=want.spectrum if p == n_parts // 2 else np.zeros(0, dtype=np.uint64),
            read_kmer_positions=want.read_kmer_positions if p == 0 else 0, path_kmer_positions=want.path_kmer_positions if p == 0 else 0,
            read_minimizers_emitted=0, path_minimizers_emitted=want.path_minimizers_emitted if p == 0 else 0, path_hits=want.path_hits if p == 0 else 0)))
    got = multi.merge_results(parts)
    whole = grouped(want)
    for f in ("rank_off", "group_len", "group_vtx", "group_member_off", "member_walk"):
        assert np.array_equal(getattr(got, f), getattr(whole, f)), f
    for f in ("spectrum", "anchor_rank", "anchor_walk", "anchor_off", "anchor_vtx", "minimizers_per_walk", "anchors_per_walk"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f
    assert got.n_filtered == want.n_filtered and got.count_sp_r == want.count_sp_r
    assert got.path_kmer_positions == want.path_kmer_positions and got.path_hits == want.path_hits


def test_merge_of_walk_parts(case):
    g, rd, want = case
    parts = [grouped(p) for p in phi_io.split_result_by_walks(want, [0, 3, 5, 9])]
    got = multi.merge_results(parts)
    for f in ("spectrum", "anchor_rank", "anchor_walk", "anchor_off", "anchor_vtx", "minimizers_per_walk", "anchors_per_walk"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("k,w", [(31, 25), (5, 3), (32, 256)])
def test_region_slices_cover_the_walks_with_context(case, world, k, w):
    g = case[0]
    seg_len = np.diff(g.seg_off.astype(np.int64))
    order = np.argsort(g.top_order_map)
    coord = np.zeros(g.n_vtx, dtype=np.int64)
    coord[order] = np.concatenate([[0], np.cumsum(seg_len[order])])[:-1]
    b = multi.region_bounds(g, world)
    assert b[0] == 0 and b[-1] == 2 ** 64 - 1 and np.all(np.diff(b.astype(np.float64)) >= 0)
    wo = g.walk_off.astype(np.int64)
    owned = [np.zeros(int(wo[h + 1] - wo[h]), dtype=np.int64) for h in range(g.n_walks)]
    for r in range(world):
        gs = multi.slice_walks(g, k, w, b[r], b[r + 1])
        assert gs.n_walks == g.n_walks
        so = gs.walk_off.astype(np.int64)
        for h in range(g.n_walks):
            full = g.walk_vtx[wo[h]:wo[h + 1]]
            sl = gs.walk_vtx[so[h]:so[h + 1]]
            c = coord[full]
            mine = (c >= int(b[r])) & (c < (int(b[r + 1]) if r + 1 < world else 2 ** 63))
            owned[h] += mine
            if not mine.any():
                assert len(sl) == 0
                continue
            a, z = np.nonzero(mine)[0][[0, -1]]
            # the slice is a contiguous piece of the walk around the owned steps ...
            first = next(i for i in range(a + 1) if np.array_equal(full[i:i + len(sl)], sl) and i + len(sl) > z)
            # ... with >= w bases in front of the first owned step and >= k-1 behind the last one, unless the walk ends there
            assert first == 0 or seg_len[full[first:a]].sum() >= w
            end = first + len(sl)
            assert end == len(full) or seg_len[full[z + 1:end]].sum() >= k - 1
    for h in range(g.n_walks):
        assert np.all(owned[h] == 1)                       # every step is owned by exactly one region


@pytest.mark.parametrize("threads", [1, 4])
def test_all_regions_at_once_equal_region_by_region(case, threads, monkeypatch):
    """phi_shard_slice_walks_all (one order check for all regions, pieces of the walks on several host threads) against one
    phi_shard_slice_walks call per region; the bounds do not depend on the number of threads either."""
    g = case[0]
    monkeypatch.setenv("PHI_SHARD_THREADS", "1")
    b1 = multi.region_bounds(g, 5)
    monkeypatch.setenv("PHI_SHARD_THREADS", str(threads))
    b = multi.region_bounds(g, 5)
    assert np.array_equal(b, b1)
    first, length = multi.slice_walks_all(g, 31, 25, b)
    wo = g.walk_off.astype(np.int64)
    for r in range(5):
        gs = multi.slice_walks(g, 31, 25, b[r], b[r + 1])
        assert np.array_equal(np.diff(gs.walk_off.astype(np.int64)), length[r].astype(np.int64))
        for h in range(g.n_walks):
            f, n = int(first[r, h]), int(length[r, h])
            assert wo[h] <= f and f + n <= wo[h + 1]
            assert np.array_equal(g.walk_vtx[f:f + n], gs.walk_vtx[int(gs.walk_off[h]):int(gs.walk_off[h + 1])])
    # one pair of steps out of order anywhere refuses the cut — also when the pair lies across a border of the pieces the parallel
    # check works on (PHI_SHARD_PIECE=4: borders every 4 steps), or at the very end of a walk
    monkeypatch.setenv("PHI_SHARD_PIECE", "4")
    assert np.array_equal(multi.region_bounds(g, 5), b1)
    f4, l4 = multi.slice_walks_all(g, 31, 25, b)
    assert np.array_equal(f4, first) and np.array_equal(l4, length)
    for s in [int(wo[3]) + d for d in (0, 2, 3, 4, 7, 8)] + [int(wo[4]) - 2, int(wo[g.n_walks]) - 2]:
        bad_vtx = g.walk_vtx.copy()
        bad_vtx[s], bad_vtx[s + 1] = bad_vtx[s + 1], bad_vtx[s]
        bad = _abi.Graph(g.seg_off, g.seg_bases, g.walk_off, bad_vtx, g.top_order_map)
        assert multi.slice_walks_all(bad, 31, 25, b) is None and multi.slice_walks(bad, 31, 25, b[0], b[1]) is None, s
    # the last step of one walk and the first of the next are not a pair
    ok_vtx = g.walk_vtx.copy()
    assert multi.slice_walks_all(_abi.Graph(g.seg_off, g.seg_bases, g.walk_off, ok_vtx, g.top_order_map), 31, 25, b) is not None


def test_split_by_weight_cuts():
    off = np.concatenate([[0], np.cumsum(np.array([5, 0, 0, 7, 1, 1, 1, 30, 2, 0], dtype=np.uint64))]).astype(np.uint64)
    for world in (1, 2, 3, 4, 7, 12):
        b = multi.split_by_weight(off, world)
        assert b[0] == 0 and b[-1] == len(off) - 1 and np.all(np.diff(b.astype(np.int64)) >= 0)
        total = int(off[-1])
        for r in range(1, world):                      # first item whose start offset reaches r/world of the total weight
            target = total * r // world
            want = next(i for i in range(len(off)) if int(off[i]) >= target)
            assert int(b[r]) == min(want, len(off) - 1), (world, r)


def test_region_cut_refuses_walks_against_the_order(case):
    g = case[0]
    bad = _abi.Graph(g.seg_off, g.seg_bases, g.walk_off, g.walk_vtx[::-1].copy(), g.top_order_map)
    assert multi.slice_walks(bad, 31, 25, 0, 1000) is None
    assert multi.region_bounds(_abi.Graph(g.seg_off, g.seg_bases, g.walk_off, g.walk_vtx, np.zeros(g.n_vtx, dtype=np.int32)), 2) is None
    gs, rs, base, region = multi.shard_inputs(bad, case[1], 1, 2, 31, 25, "region")    # falls back to whole walks
    assert region is None and base > 0


@pytest.mark.parametrize("world", [2, 3])
def test_bench_generates_the_same_shards_it_would_cut_from_the_whole_input(monkeypatch, world):
    """bench.py never spells the whole walk set of the chromosome-scale configs: a rank generates only the part of every walk that
    runs through its region, and only its batch range of the reads.  On a small instance of the same generator that must be exactly
    what phi_shard_walk_regions / phi_shard_slice_walks cut from the whole graph, and the read shards must tile the global read set."""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py", "--config", "c4", "--haps", "11", "--backbone", "300000", "--coverage", "2"])
    a = bench.parse_args()
    whole_g, whole_rd, _, _, _, units = bench.Workload(a).shard(0, 1)
    assert units[0] == bench.positions(whole_g.walk_lengths(), a.k, a.w)
    # the region bounds bench derives from 4 sample walks
    sample = [h for h in range(0, a.haps, max(1, a.haps // 4))][:4]
    wo = whole_g.walk_off.astype(np.int64)
    gsample = _abi.Graph(whole_g.seg_off, whole_g.seg_bases, np.concatenate([[0], np.cumsum([wo[h + 1] - wo[h] for h in sample])]),
                         np.concatenate([whole_g.walk_vtx[wo[h]:wo[h + 1]] for h in sample]), whole_g.top_order_map)
    bounds = multi.region_bounds(gsample, world)
    reads = []
    for r in range(world):
        g, rd, base, n_walks, region, units_r = bench.Workload(a).shard(r, world)
        assert (base, n_walks, units_r) == (0, a.haps, units) and region == (int(bounds[r]), int(bounds[r + 1]))
        want = multi.slice_walks(whole_g, a.k, a.w, bounds[r], bounds[r + 1])
        assert np.array_equal(g.walk_off, want.walk_off) and np.array_equal(g.walk_vtx, want.walk_vtx)
        reads.append(rd)
    assert sum(r.n_reads for r in reads) == whole_rd.n_reads
    assert np.array_equal(np.concatenate([r.read_bases for r in reads]), whole_rd.read_bases)


@pytest.mark.parametrize("threads", [1, 4])
def test_merge_of_empty_parts(monkeypatch, threads):
    """No reads (count_sp_r == 0) on every GPU: the merge is a result without ranks whose per-walk counters are still the sums."""
    monkeypatch.setenv("PHI_MERGE_THREADS", str(threads))
    z = np.zeros(0)

    def empty(nw):
        return _abi.IndexResultPy(count_sp_r=0, n_walks=nw, n_filtered=0, spectrum=z.astype(np.uint64), anchor_rank=z.astype(np.int32),
                                  anchor_walk=z.astype(np.int32), anchor_off=np.zeros(1, dtype=np.uint64), anchor_vtx=z.astype(np.int32),
                                  minimizers_per_walk=np.arange(nw, dtype=np.uint64), anchors_per_walk=np.zeros(nw, dtype=np.uint64),
                                  n_groups=0, group_len=z.astype(np.uint8), group_vtx=z.astype(np.int32),
                                  group_member_off=np.zeros(1, dtype=np.uint32), member_walk=z.astype(np.int32), rank_off=np.zeros(1, dtype=np.uint32))
    for n_parts in (1, 3):
        got = multi.merge_results([empty(4) for _ in range(n_parts)])
        assert got.count_sp_r == 0 and got.n_groups == 0 and got.n_anchors == 0
        assert got.minimizers_per_walk.tolist() == [0, n_parts, 2 * n_parts, 3 * n_parts]
