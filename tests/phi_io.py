"""Test helpers: .phiarr reader (written by oracle/ref_build/ref_probe.cpp), model-dump parser
(written by oracle/ref_build/stub/gurobi_c++.h), oracle ctypes binding.  Test infrastructure only."""
import ctypes as C
import hashlib
import os
import re
import struct
import subprocess

import numpy as np

from phi_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

_DT = {"B": np.uint8, "i": np.int32, "I": np.uint32, "Q": np.uint64, "q": np.int64}


def read_phiarr(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    p = 0
    while p < len(data):
        (nl,) = struct.unpack_from("<I", data, p); p += 4
        name = data[p:p + nl].decode(); p += nl
        dt = chr(data[p]); p += 1
        (n,) = struct.unpack_from("<Q", data, p); p += 8
        dtype = np.dtype(_DT[dt])
        out[name] = np.frombuffer(data, dtype=dtype, count=n, offset=p).copy()
        p += n * dtype.itemsize
    return out


def graph_from_arrays(d):
    names = bytes(d["walk_names"]).decode().split("\n")[:-1] if "walk_names" in d else []
    return _abi.Graph(d["seg_off"], d["seg_bases"], d["walk_off"], d["walk_vtx"], d["top_order_map"], names)


def reads_from_arrays(d):
    return _abi.Reads(d["read_off"], d["read_bases"])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------- model dump
_EDGE = re.compile(r"^(\d+)_(\d+)_(\d+)_(\d+)$")


def parse_model_dump(path):
    """Recover what the dump pins about Anchor_hits: for every z_i_j_k (creation order) the vertex list
    implied by its edge variables (None for single-vertex anchors, which create no constraint terms).
    Works for both -q1 (Q records) and -q0 (C Kmer_constraints_i_j_k records)."""
    z_order = []          # (i, j, k) in creation order
    lists = {}            # (i, j, k) -> [v0, v1, ...]
    n = {"V": 0, "C": 0, "Q": 0}
    with open(path) as f:
        for line in f:
            t = line[0]
            if t in n:
                n[t] += 1
            if t == "V" and line.startswith("V z_"):
                parts = line.split()[1].split("_")
                if len(parts) == 4:
                    z_order.append((int(parts[1]), int(parts[2]), int(parts[3])))
            elif t == "Q" and line.startswith("Q Kmer_constraints_"):
                quad = line.split("|")[1].split(";")[1].split()
                cur = {}
                for term in quad:               # coeff*u_j_v_j*z_i_j_k, in walk order
                    _, e, z = term.split("*")
                    m = _EDGE.match(e)
                    zi = tuple(int(x) for x in z.split("_")[1:])
                    lst = cur.setdefault(zi, [])
                    if not lst:
                        lst.append(int(m.group(1)))
                    lst.append(int(m.group(3)))
                lists.update(cur)
            elif t == "C" and line.startswith("C Kmer_constraints_"):
                name = line.split()[1]
                idx = name[len("Kmer_constraints_"):].split("_")
                if len(idx) == 3:
                    lhs = line.split("|")[1].split()
                    lst = []
                    for term in lhs:
                        _, e = term.split("*")
                        m = _EDGE.match(e)
                        if not lst:
                            lst.append(int(m.group(1)))
                        lst.append(int(m.group(3)))
                    lists[tuple(int(x) for x in idx)] = lst
    return z_order, lists, n


# -------------------------------------------------------------------- oracle
_oracle = None


class OracleResult(C.Structure):
    """oracle/phi_oracle.h: phi_oracle_result (the reference's final Anchor_hits, anchor by anchor)."""
    _fields_ = [("count_sp_r", C.c_int32), ("n_walks", C.c_uint32), ("n_filtered", C.c_int64),
                ("n_anchors", C.c_uint64), ("n_anchor_vtx", C.c_uint64),
                ("spectrum", _abi.u64p), ("rank_off", _abi.u64p), ("anchor_walk", _abi.i32p), ("anchor_len", _abi.u8p),
                ("anchor_vtx", _abi.i32p), ("minimizers_per_walk", _abi.u64p), ("anchors_per_walk", _abi.u64p),
                ("read_kmer_positions", C.c_uint64), ("path_kmer_positions", C.c_uint64),
                ("read_minimizers_emitted", C.c_uint64), ("path_minimizers_emitted", C.c_uint64),
                ("path_hits", C.c_uint64), ("n_walk_kmers", C.c_uint64), ("shared_kmer_hist", _abi.u64p)]


def oracle_result_to_py(res):
    """phi_oracle_result -> the same IndexResultPy the product's results are expanded into (anchor_* fields)."""
    na, nv, nw, ns = res.n_anchors, res.n_anchor_vtx, res.n_walks, res.count_sp_r
    rank_off = _abi._np_from(res.rank_off, ns + 1 if res.rank_off else 0, np.uint64)
    lens = _abi._np_from(res.anchor_len, na, np.uint8)
    if len(rank_off):
        assert int(rank_off[-1]) == na and int(rank_off[0]) == 0
        anchor_rank = np.repeat(np.arange(ns, dtype=np.int32), np.diff(rank_off.astype(np.int64)))
    else:
        anchor_rank = np.zeros(na, dtype=np.int32)
    anchor_off = np.concatenate([[0], np.cumsum(lens, dtype=np.uint64)]).astype(np.uint64)
    return _abi.IndexResultPy(
        count_sp_r=int(ns), n_walks=int(nw), n_filtered=int(res.n_filtered),
        spectrum=_abi._np_from(res.spectrum, ns, np.uint64),
        anchor_rank=anchor_rank, anchor_walk=_abi._np_from(res.anchor_walk, na, np.int32),
        anchor_off=anchor_off, anchor_vtx=_abi._np_from(res.anchor_vtx, nv, np.int32),
        minimizers_per_walk=_abi._np_from(res.minimizers_per_walk, nw, np.uint64),
        anchors_per_walk=_abi._np_from(res.anchors_per_walk, nw, np.uint64),
        read_kmer_positions=int(res.read_kmer_positions), path_kmer_positions=int(res.path_kmer_positions),
        read_minimizers_emitted=int(res.read_minimizers_emitted),
        path_minimizers_emitted=int(res.path_minimizers_emitted), path_hits=int(res.path_hits),
        n_walk_kmers=int(res.n_walk_kmers),
        shared_kmer_hist=_abi._np_from(res.shared_kmer_hist, nw + 1, np.uint64) if res.shared_kmer_hist else None)


def oracle_lib():
    """ctypes handle on oracle/libphi_oracle.so, building it with gcc if needed (test infra only)."""
    global _oracle
    if _oracle is None:
        so = os.path.join(ROOT, "oracle", "libphi_oracle.so")
        src = os.path.join(ROOT, "oracle", "phi_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libphi_oracle.so"],
                                  stdout=subprocess.DEVNULL)
        lib = C.CDLL(so)
        lib.phi_oracle_hash128_to_64.restype = C.c_uint64
        lib.phi_oracle_hash128_to_64.argtypes = [C.c_char_p, C.c_int32]
        lib.phi_oracle_index_run.restype = C.c_int
        lib.phi_oracle_index_run.argtypes = [C.POINTER(_abi.GraphView), C.POINTER(_abi.ReadsView),
                                             C.POINTER(_abi.IndexParams), C.c_int,
                                             C.POINTER(C.POINTER(OracleResult))]
        lib.phi_oracle_sketch_walks.restype = C.c_int
        lib.phi_oracle_sketch_walks.argtypes = [C.POINTER(_abi.GraphView), C.POINTER(_abi.IndexParams), C.c_int,
                                                C.POINTER(C.POINTER(OracleResult)), C.POINTER(_abi.u64p)]
        lib.phi_oracle_sketch_walks_owner.restype = C.c_int
        lib.phi_oracle_sketch_walks_owner.argtypes = [C.POINTER(_abi.GraphView), C.POINTER(_abi.IndexParams), C.c_int,
                                                      C.POINTER(C.POINTER(OracleResult)), C.POINTER(_abi.u64p), C.POINTER(_abi.i32p)]
        lib.phi_oracle_read_hashes.restype = C.c_int64
        lib.phi_oracle_read_hashes.argtypes = [C.c_char_p, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(_abi.u64p)]
        lib.phi_oracle_result_free.argtypes = [C.POINTER(OracleResult)]
        lib.phi_oracle_free.argtypes = [C.c_void_p]
        _oracle = lib
    return _oracle


def oracle_hash(key: bytes) -> int:
    return oracle_lib().phi_oracle_hash128_to_64(key, len(key))


def oracle_index(graph, reads, k=31, w=25, threshold=1.0, threads=0, debug=0):
    lib = oracle_lib()
    gv, rv = graph.view(), reads.view()
    prm = _abi.IndexParams(k, w, threshold, debug)
    out = C.POINTER(OracleResult)()
    rc = lib.phi_oracle_index_run(C.byref(gv), C.byref(rv), C.byref(prm), threads, C.byref(out))
    assert rc == 0, rc
    res = oracle_result_to_py(out.contents)
    lib.phi_oracle_result_free(out)
    return res


def oracle_sketch_walks(graph, k=31, w=25, threads=0):
    lib = oracle_lib()
    gv = graph.view()
    prm = _abi.IndexParams(k, w, 1.0, 0)
    out = C.POINTER(OracleResult)()
    hp = _abi.u64p()
    rc = lib.phi_oracle_sketch_walks(C.byref(gv), C.byref(prm), threads, C.byref(out), C.byref(hp))
    assert rc == 0, rc
    res = oracle_result_to_py(out.contents)
    hashes = _abi._np_from(hp, res.n_anchors, np.uint64)
    lib.phi_oracle_result_free(out)
    lib.phi_oracle_free(hp)
    return res, hashes


def oracle_sketch_walks_owner(graph, k=31, w=25, threads=0):
    """oracle_sketch_walks + for every emitted minimizer the vertex under the start of the last k-mer of the window that emitted it."""
    lib = oracle_lib()
    gv = graph.view()
    prm = _abi.IndexParams(k, w, 1.0, 0)
    out = C.POINTER(OracleResult)()
    hp, op = _abi.u64p(), _abi.i32p()
    rc = lib.phi_oracle_sketch_walks_owner(C.byref(gv), C.byref(prm), threads, C.byref(out), C.byref(hp), C.byref(op))
    assert rc == 0, rc
    res = oracle_result_to_py(out.contents)
    hashes = _abi._np_from(hp, res.n_anchors, np.uint64)
    owner = _abi._np_from(op, res.n_anchors, np.int32)
    lib.phi_oracle_result_free(out)
    lib.phi_oracle_free(hp)
    lib.phi_oracle_free(op)
    return res, hashes, owner


def oracle_read_hashes(seq: bytes, k=31, w=25):
    lib = oracle_lib()
    hp = _abi.u64p()
    n = lib.phi_oracle_read_hashes(seq, len(seq), k, w, C.byref(hp))
    a = _abi._np_from(hp, n, np.uint64)
    lib.phi_oracle_free(hp)
    return a


# ---- the grouped result form (include/phi_gpu_index.h, ABI v3) derived from per-anchor results: test infrastructure for the
# CPU tests of integration/phi_model.hpp
def group_anchors(res):
    """IndexResultPy with anchors in (rank, walk, j) order -> (rank_off u32, group_len u8, group_vtx i32, group_member_off u32,
    member_walk i32): per rank the distinct vertex lists in std::map<std::string> order of "v0_v1_..._"
    (/root/reference/src/ILP_index.cpp:680-709), per list the walks that carry it, ascending."""
    off = res.anchor_off.astype(np.int64)
    rank_off = np.zeros(res.count_sp_r + 1, dtype=np.uint32)
    glen, gvtx, moff, mwalk = [], [], [0], []
    a, na = 0, res.n_anchors
    for r in range(res.count_sp_r):
        groups = {}
        while a < na and int(res.anchor_rank[a]) == r:
            groups.setdefault(tuple(res.anchor_vtx[off[a]:off[a + 1]].tolist()), []).append(int(res.anchor_walk[a]))
            a += 1
        for key in sorted(groups, key=lambda t: "".join(f"{v}_" for v in t).encode()):
            glen.append(len(key)); gvtx.extend(key)
            mwalk.extend(sorted(groups[key])); moff.append(len(mwalk))
        rank_off[r + 1] = len(glen)
    return (rank_off, np.array(glen, dtype=np.uint8), np.array(gvtx, dtype=np.int32), np.array(moff, dtype=np.uint32),
            np.array(mwalk, dtype=np.int32))


def write_result_file(path, res, member_bytes=2):
    """integration/phi_adapter_testhook.hpp's file format, from a per-anchor result (the oracle's)."""
    rank_off, glen, gvtx, moff, mwalk = group_anchors(res)
    head = np.array([res.count_sp_r, res.n_walks, res.n_filtered, len(mwalk), len(glen), len(gvtx), member_bytes], dtype=np.uint64)
    arrays = [res.spectrum.astype(np.uint64), rank_off, glen, gvtx, moff,
              mwalk.astype(np.uint16) if member_bytes == 2 else mwalk, res.minimizers_per_walk.astype(np.uint64),
              res.anchors_per_walk.astype(np.uint64)]
    with open(path, "wb") as f:
        f.write(b"PHIRES3\0")
        f.write(head.tobytes())
        for a in arrays:
            b = a.tobytes()
            f.write(b + b"\0" * (-len(b) % 8))


def split_result_by_walks(res, bounds):
    """Per-anchor result -> what every GPU of a multi-GPU run returns (the anchors of ITS walks for all ranks; per-walk counters are
    partial sums; n_filtered is reported once): test infrastructure for the multi-part adapter paths."""
    import dataclasses
    parts = []
    lens = np.diff(res.anchor_off.astype(np.int64))
    for p in range(len(bounds) - 1):
        lo, hi = int(bounds[p]), int(bounds[p + 1])
        keep = (res.anchor_walk >= lo) & (res.anchor_walk < hi)
        idx = np.nonzero(keep)[0]
        starts = res.anchor_off.astype(np.int64)[idx]
        vtx = np.concatenate([res.anchor_vtx[s:s + n] for s, n in zip(starts, lens[idx])]) if len(idx) else np.zeros(0, dtype=np.int32)
        in_range = (np.arange(res.n_walks) >= lo) & (np.arange(res.n_walks) < hi)
        parts.append(dataclasses.replace(
            res, n_filtered=res.n_filtered if p == 0 else 0,
            anchor_rank=res.anchor_rank[idx], anchor_walk=res.anchor_walk[idx],
            anchor_off=np.concatenate([[0], np.cumsum(lens[idx])]).astype(np.uint64), anchor_vtx=vtx.astype(np.int32),
            minimizers_per_walk=np.where(in_range, res.minimizers_per_walk, 0).astype(np.uint64),
            anchors_per_walk=np.where(in_range, res.anchors_per_walk, 0).astype(np.uint64)))
    return parts
