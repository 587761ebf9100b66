"""Host-side cost of phi_index_result_merge on parts shaped like configs[3] (c4) split over N GPUs by region: 8.85 M hash ranks,
1.59 M surviving groups, 204 M anchors (u16 member walks), 6 M group vertices; every rank's groups come from one part except a
few per cent whose groups are split over two parts.  CPU only:  python profiles/merge_bench.py [n_parts] [scale]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phi_b200 import _abi
from phi_b200.api import load_library


def make_whole(rng, ns, n_walks=200):
    has = rng.random(ns) < 0.14
    per_rank = np.where(has, 1 + (rng.random(ns) < 0.22) + (rng.random(ns) < 0.06), 0).astype(np.int64)
    rank_off = np.concatenate([[0], np.cumsum(per_rank)]).astype(np.uint32)
    ng = int(rank_off[-1])
    g_rank = np.repeat(np.arange(ns), per_rank)
    g_in_rank = np.arange(ng) - rank_off[g_rank].astype(np.int64)
    glen = rng.integers(2, 6, ng).astype(np.uint8)
    voff = np.concatenate([[0], np.cumsum(glen.astype(np.int64))])
    gvtx = rng.integers(1000000, 9999999, int(voff[-1])).astype(np.int32)
    gvtx[voff[:-1]] = 1000000 + g_in_rank                                 # key order inside a rank = group order
    cnt = rng.integers(1, 2 * 128, ng).astype(np.int64)
    cnt = np.minimum(cnt, n_walks)
    start = (rng.random(ng) * (n_walks - cnt + 1)).astype(np.int64)
    moff = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint32)
    members = (np.arange(int(moff[-1])) - np.repeat(moff[:-1].astype(np.int64), cnt) + np.repeat(start, cnt)).astype(np.uint16)
    return dict(ns=ns, n_walks=n_walks, rank_off=rank_off, glen=glen, gvtx=gvtx, voff=voff, moff=moff, members=members, g_rank=g_rank, per_rank=per_rank)


def take_groups(w, sel):
    """The part that holds the groups sel (bool per group), as ABI arrays."""
    idx = np.nonzero(sel)[0]
    per_rank = np.bincount(w["g_rank"][idx], minlength=w["ns"])
    rank_off = np.concatenate([[0], np.cumsum(per_rank)]).astype(np.uint32)
    glen = w["glen"][idx]
    vsel = np.repeat(sel, w["glen"].astype(np.int64))
    cnt = np.diff(w["moff"].astype(np.int64))
    msel = np.repeat(sel, cnt)
    moff = np.concatenate([[0], np.cumsum(cnt[idx])]).astype(np.uint32)
    return dict(rank_off=rank_off, glen=np.ascontiguousarray(glen), gvtx=np.ascontiguousarray(w["gvtx"][vsel]), moff=moff, members=np.ascontiguousarray(w["members"][msel]))


def c_part(w, p, first):
    r = _abi.IndexResult()
    r.count_sp_r, r.n_walks = w["ns"], w["n_walks"]
    r.n_anchors, r.n_groups, r.n_group_vtx = len(p["members"]), len(p["glen"]), len(p["gvtx"])
    r.rank_off = p["rank_off"].ctypes.data_as(_abi.u32p)
    r.group_len = p["glen"].ctypes.data_as(_abi.u8p)
    r.group_vtx = p["gvtx"].ctypes.data_as(_abi.i32p)
    r.group_member_off = p["moff"].ctypes.data_as(_abi.u32p)
    r.member_walk16 = p["members"].ctypes.data_as(_abi.u16p)
    zeros = np.zeros(w["n_walks"], dtype=np.uint64)
    p["_z"] = zeros
    r.minimizers_per_walk = zeros.ctypes.data_as(_abi.u64p)
    r.anchors_per_walk = zeros.ctypes.data_as(_abi.u64p)
    if first:
        p["_s"] = np.arange(w["ns"], dtype=np.uint64)
        r.spectrum = p["_s"].ctypes.data_as(_abi.u64p)
    return r


def main():
    n_parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    rng = np.random.default_rng(7)
    w = make_whole(rng, int(8852414 * scale))
    ng = len(w["glen"])
    owner = rng.integers(0, n_parts, w["ns"])[w["g_rank"]]                # all groups of a rank on one part ...
    split = (rng.random(w["ns"]) < 0.03)[w["g_rank"]] & (w["per_rank"][w["g_rank"]] > 1)
    g_in_rank = np.arange(ng) - w["rank_off"][w["g_rank"]].astype(np.int64)
    owner = np.where(split & (g_in_rank % 2 == 1), (owner + 1) % n_parts, owner)   # ... but 3 % of the multi-group ranks are split over two
    parts = [take_groups(w, owner == p) for p in range(n_parts)]
    cs = [c_part(w, p, i == 0) for i, p in enumerate(parts)]
    lib = load_library()
    arr = (C.POINTER(_abi.IndexResult) * n_parts)(*[C.pointer(c) for c in cs])
    print(f"parts {n_parts}: ranks {w['ns']}, groups {ng}, anchors {len(w['members'])}, group vertices {len(w['gvtx'])}, "
          f"input {sum(sum(a.nbytes for a in p.values()) for p in parts) / 1e6:.0f} MB")
    for threads in ([1, 2, 4, 8, 16] if "PHI_MERGE_THREADS" not in os.environ else [int(os.environ["PHI_MERGE_THREADS"])]):
        os.environ["PHI_MERGE_THREADS"] = str(threads)
        best = 1e9
        for rep in range(5):
            out = C.POINTER(_abi.IndexResult)()
            t0 = time.perf_counter()
            rc = lib.phi_index_result_merge(arr, n_parts, C.byref(out))
            dt = time.perf_counter() - t0
            assert rc == 0
            best = min(best, dt)
            if rep == 0:
                o = out.contents
                assert o.n_groups == ng and o.n_anchors == len(w["members"]) and o.n_group_vtx == len(w["gvtx"])
                assert np.array_equal(np.ctypeslib.as_array(o.rank_off, (w["ns"] + 1,)), w["rank_off"])
                assert np.array_equal(np.ctypeslib.as_array(o.group_len, (ng,)), w["glen"])
                assert np.array_equal(np.ctypeslib.as_array(o.group_vtx, (len(w["gvtx"]),)), w["gvtx"])
                assert np.array_equal(np.ctypeslib.as_array(o.group_member_off, (ng + 1,)), w["moff"])
                assert np.array_equal(np.ctypeslib.as_array(o.member_walk16, (len(w["members"]),)), w["members"])
                assert np.array_equal(np.ctypeslib.as_array(o.spectrum, (w["ns"],)), parts[0]["_s"])
            lib.phi_gpu_index_result_free(out)
        print(f"  threads {threads:2d}: {best * 1e3:8.1f} ms (best of 5; identical to the whole: yes)")


if __name__ == "__main__":
    main()
