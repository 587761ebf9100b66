"""ctypes mirror of include/phi_gpu_index.h (plain C structs, no torch types).

The same structs describe the product library (phi_b200/libphi_gpu_index.so)
and, in tests only, the CPU oracle (oracle/libphi_oracle.so).
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

PHI_OK = 0
PHI_ERR_ARG, PHI_ERR_UNSUPPORTED, PHI_ERR_CUDA, PHI_ERR_NOMEM, PHI_ERR_COMM = 1, 2, 3, 4, 5
PHI_COMM_ID_BYTES = 128

u8p, u32p, i32p, u64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_uint64))
u16p = C.POINTER(C.c_uint16)


class GraphView(C.Structure):
    _fields_ = [("n_vtx", C.c_uint32), ("seg_off", u64p), ("seg_bases", u8p), ("n_walks", C.c_uint32),
                ("walk_off", u64p), ("walk_vtx", u32p), ("top_order_map", i32p)]


class ReadsView(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("read_off", u64p), ("read_bases", u8p)]


class IndexParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("w", C.c_int32), ("threshold", C.c_float), ("debug", C.c_int32)]


class IndexResult(C.Structure):
    _fields_ = [("count_sp_r", C.c_int32), ("n_walks", C.c_uint32), ("n_filtered", C.c_int64),
                ("n_anchors", C.c_uint64), ("n_groups", C.c_uint64), ("n_group_vtx", C.c_uint64),
                ("spectrum", u64p), ("rank_off", u32p), ("group_len", u8p), ("group_vtx", i32p),
                ("group_member_off", u32p), ("member_walk16", u16p), ("member_walk32", i32p),
                ("minimizers_per_walk", u64p), ("anchors_per_walk", u64p),
                ("read_kmer_positions", C.c_uint64), ("path_kmer_positions", C.c_uint64),
                ("read_minimizers_emitted", C.c_uint64), ("path_minimizers_emitted", C.c_uint64),
                ("path_hits", C.c_uint64), ("n_walk_kmers", C.c_uint64), ("shared_kmer_hist", u64p)]


class StageTimes(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("graph_prep_ms", C.c_float), ("read_sketch_ms", C.c_float),
                ("spectrum_ms", C.c_float), ("walk_sketch_ms", C.c_float), ("filter_ms", C.c_float),
                ("d2h_ms", C.c_float), ("total_ms", C.c_float), ("walk_kernel_ms", C.c_float),
                ("read_kernel_ms", C.c_float), ("kernel_launches", C.c_uint64),
                ("exchange_spectrum_ms", C.c_float), ("route_hits_ms", C.c_float), ("exchange_hits_ms", C.c_float)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SharingStats(C.Structure):
    _fields_ = [("chunks", C.c_uint64), ("active_chunks", C.c_uint64), ("tiles", C.c_uint64),
                ("unique_windows", C.c_uint64), ("unique_hits", C.c_uint64)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def _arr(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


@dataclass
class Graph:
    """Flat graph view; see phi_graph_view in include/phi_gpu_index.h."""
    seg_off: np.ndarray
    seg_bases: np.ndarray
    walk_off: np.ndarray
    walk_vtx: np.ndarray
    top_order_map: np.ndarray
    walk_names: list = field(default_factory=list)

    def __post_init__(self):
        self.seg_off = _arr(self.seg_off, np.uint64)
        self.seg_bases = _arr(self.seg_bases, np.uint8)
        self.walk_off = _arr(self.walk_off, np.uint64)
        self.walk_vtx = _arr(self.walk_vtx, np.uint32)
        self.top_order_map = _arr(self.top_order_map, np.int32)

    @property
    def n_vtx(self):
        return len(self.seg_off) - 1

    @property
    def n_walks(self):
        return len(self.walk_off) - 1

    def view(self):
        return GraphView(self.n_vtx, self.seg_off.ctypes.data_as(u64p), self.seg_bases.ctypes.data_as(u8p),
                         self.n_walks, self.walk_off.ctypes.data_as(u64p), self.walk_vtx.ctypes.data_as(u32p),
                         self.top_order_map.ctypes.data_as(i32p))

    def walk_lengths(self):
        seg_len = np.diff(self.seg_off.astype(np.int64))
        step_len = seg_len[self.walk_vtx]
        cs = np.concatenate([[0], np.cumsum(step_len)])
        return cs[self.walk_off.astype(np.int64)[1:]] - cs[self.walk_off.astype(np.int64)[:-1]]

    def take_walks(self, lo, hi):
        """Sub-graph view holding walks [lo, hi) (segments replicated)."""
        wo = self.walk_off.astype(np.int64)
        return Graph(self.seg_off, self.seg_bases, wo[lo:hi + 1] - wo[lo], self.walk_vtx[wo[lo]:wo[hi]],
                     self.top_order_map, self.walk_names[lo:hi])


@dataclass
class Reads:
    read_off: np.ndarray
    read_bases: np.ndarray

    def __post_init__(self):
        self.read_off = _arr(self.read_off, np.uint64)
        self.read_bases = _arr(self.read_bases, np.uint8)

    @property
    def n_reads(self):
        return len(self.read_off) - 1

    def view(self):
        return ReadsView(self.n_reads, self.read_off.ctypes.data_as(u64p), self.read_bases.ctypes.data_as(u8p))

    @staticmethod
    def from_strings(seqs):
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        if seqs:
            off[1:] = np.cumsum([len(s) for s in seqs])
        data = "".join(seqs).encode("latin-1") if seqs and isinstance(seqs[0], str) else b"".join(seqs)
        return Reads(off, np.frombuffer(data, dtype=np.uint8).copy())

    def take(self, lo, hi):
        ro = self.read_off.astype(np.int64)
        return Reads(ro[lo:hi + 1] - ro[lo], self.read_bases[ro[lo]:ro[hi]])


@dataclass
class IndexResultPy:
    """Host copy of phi_index_result (numpy arrays).  group_* / member_walk are the ABI's arrays (the reference's filter map:
    one vertex list per group plus the walks that carry it); anchor_rank / anchor_walk / anchor_off / anchor_vtx are what the
    reference-side adapter's forward pass builds from them: one entry per anchor in (rank, walk, j) order, i.e.
    Anchor_hits[rank][walk][j]."""
    count_sp_r: int
    n_walks: int
    n_filtered: int
    spectrum: np.ndarray
    anchor_rank: np.ndarray
    anchor_walk: np.ndarray
    anchor_off: np.ndarray
    anchor_vtx: np.ndarray
    minimizers_per_walk: np.ndarray
    anchors_per_walk: np.ndarray
    read_kmer_positions: int = 0
    path_kmer_positions: int = 0
    read_minimizers_emitted: int = 0
    path_minimizers_emitted: int = 0
    path_hits: int = 0
    n_walk_kmers: int = 0
    shared_kmer_hist: np.ndarray = None        # [n_walks + 1] when the run had debug != 0
    n_groups: int = 0
    group_rank: np.ndarray = None
    group_len: np.ndarray = None
    group_vtx: np.ndarray = None
    group_member_off: np.ndarray = None
    member_walk: np.ndarray = None
    member_walk_bytes: int = 4
    rank_off: np.ndarray = None                # [count_sp_r + 1] first group of every rank (the ABI's array)

    @property
    def n_anchors(self):
        return len(self.anchor_rank) if len(self.anchor_rank) or self.member_walk is None else len(self.member_walk)

    def wire_bytes(self):
        """Bytes of the C result arrays (what crosses PCIe): spectrum, rank_off, group_len, group_vtx, group_member_off,
        member_walk, per-walk counters."""
        ns, ng = self.count_sp_r, self.n_groups
        return 8 * len(self.spectrum) + 4 * (ns + 1) + ng + 4 * len(self.group_vtx) + 4 * (ng + 1) + self.member_walk_bytes * len(self.member_walk) + 16 * self.n_walks

    def anchors(self):
        """[(rank, walk, [vertices])] in final order."""
        off = self.anchor_off
        return [(int(self.anchor_rank[a]), int(self.anchor_walk[a]), self.anchor_vtx[off[a]:off[a + 1]].tolist())
                for a in range(self.n_anchors)]


def _np_from(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


def expand_groups(group_rank, group_len, group_vtx, group_member_off, member_walk):
    """The adapter's forward pass (ILP_index.cpp:700-709) in numpy: for every group in order, for every member walk h, push the
    group's vertex list onto Anchor_hits[rank][h].  Returns the anchors in (rank, walk, j) order:
    (anchor_rank, anchor_walk, anchor_off, anchor_vtx)."""
    ng = len(group_len)
    lens64 = group_len.astype(np.int64)
    gvoff = np.concatenate([[0], np.cumsum(lens64)])
    g_of_m = np.repeat(np.arange(ng, dtype=np.int64), np.diff(group_member_off.astype(np.int64)))
    m_rank = group_rank.astype(np.int64)[g_of_m]
    idx = np.lexsort((member_walk, m_rank))                  # stable: pushes onto one (rank, walk) keep the group order
    g_sorted = g_of_m[idx]
    alen = lens64[g_sorted]
    anchor_off = np.concatenate([[0], np.cumsum(alen)]).astype(np.uint64)
    total = int(anchor_off[-1])
    src = np.repeat(gvoff[g_sorted], alen) + (np.arange(total, dtype=np.int64) - np.repeat(anchor_off[:-1].astype(np.int64), alen))
    return (m_rank[idx].astype(np.int32), member_walk[idx].astype(np.int32), anchor_off,
            group_vtx[src].astype(np.int32) if total else np.zeros(0, dtype=np.int32))


def py_to_c_result(py: IndexResultPy):
    """IndexResultPy (grouped arrays) -> (IndexResult, keep-alive list): input of phi_index_result_merge."""
    rank_off = np.ascontiguousarray(py.rank_off, dtype=np.uint32)
    glen = np.ascontiguousarray(py.group_len, dtype=np.uint8)
    gvtx = np.ascontiguousarray(py.group_vtx, dtype=np.int32)
    moff = np.ascontiguousarray(py.group_member_off, dtype=np.uint32)
    w16 = py.n_walks <= 65536
    mw = np.ascontiguousarray(py.member_walk, dtype=np.uint16 if w16 else np.int32)
    spec = np.ascontiguousarray(py.spectrum, dtype=np.uint64)
    mpw = np.ascontiguousarray(py.minimizers_per_walk, dtype=np.uint64)
    apw = np.ascontiguousarray(py.anchors_per_walk, dtype=np.uint64)
    keep = [rank_off, glen, gvtx, moff, mw, spec, mpw, apw]
    r = IndexResult()
    r.count_sp_r, r.n_walks, r.n_filtered = py.count_sp_r, py.n_walks, py.n_filtered
    r.n_anchors, r.n_groups, r.n_group_vtx = len(mw), len(glen), len(gvtx)
    r.spectrum = spec.ctypes.data_as(u64p) if len(spec) == py.count_sp_r and py.count_sp_r else None
    r.rank_off = rank_off.ctypes.data_as(u32p)
    r.group_len = glen.ctypes.data_as(u8p)
    r.group_vtx = gvtx.ctypes.data_as(i32p)
    r.group_member_off = moff.ctypes.data_as(u32p)
    if w16:
        r.member_walk16 = mw.ctypes.data_as(u16p)
    else:
        r.member_walk32 = mw.ctypes.data_as(i32p)
    r.minimizers_per_walk = mpw.ctypes.data_as(u64p)
    r.anchors_per_walk = apw.ctypes.data_as(u64p)
    if py.shared_kmer_hist is not None:
        hist = np.ascontiguousarray(py.shared_kmer_hist, dtype=np.uint64)
        keep.append(hist)
        r.shared_kmer_hist = hist.ctypes.data_as(u64p)
        r.n_walk_kmers = py.n_walk_kmers
    r.read_kmer_positions, r.path_kmer_positions = py.read_kmer_positions, py.path_kmer_positions
    r.read_minimizers_emitted, r.path_minimizers_emitted, r.path_hits = py.read_minimizers_emitted, py.path_minimizers_emitted, py.path_hits
    return r, keep


def result_to_py(res: IndexResult, expand=True) -> IndexResultPy:
    """expand=False skips the adapter's forward pass (anchor_* stay empty): for big results of which only the ABI arrays are wanted."""
    na, ng, nv, nw, ns = res.n_anchors, res.n_groups, res.n_group_vtx, res.n_walks, res.count_sp_r
    have = bool(res.group_len) or ng == 0
    rank_off = _np_from(res.rank_off, ns + 1 if res.rank_off else 0, np.uint32)
    group_len = _np_from(res.group_len, ng, np.uint8)
    group_vtx = _np_from(res.group_vtx, nv, np.int32)
    walk_bytes = 2 if res.member_walk16 else 4
    member_walk = _np_from(res.member_walk16, na, np.uint16).astype(np.int32) if res.member_walk16 else _np_from(res.member_walk32, na, np.int32)
    if not have:                                               # counters only (run_resident without download)
        ng = 0
    if res.group_member_off:
        member_off = _np_from(res.group_member_off, ng + 1, np.uint32)
        if ng == 0:
            member_off = np.zeros(1, dtype=np.uint32)
    else:                                                      # sketch-only result: one member per group
        assert na == ng or not have
        member_off = np.arange(ng + 1, dtype=np.uint32)
    if have:
        assert int(member_off[0]) == 0 and int(member_off[-1]) == len(member_walk), "group_member_off does not cover member_walk"
        assert int(group_len.astype(np.int64).sum()) == len(group_vtx), "group_len does not cover group_vtx"
    if len(rank_off):
        assert int(rank_off[-1]) == ng and int(rank_off[0]) == 0
        group_rank = np.repeat(np.arange(ns, dtype=np.int32), np.diff(rank_off.astype(np.int64)))
        # members of a group ascend (the ABI promises it)
        if len(member_walk) > 1:
            same = np.ones(len(member_walk) - 1, dtype=bool)
            starts = member_off[1:-1].astype(np.int64)
            same[starts[(starts > 0) & (starts < len(member_walk))] - 1] = False
            assert np.all((np.diff(member_walk.astype(np.int64)) >= 0) | ~same), "members of a group are not ascending"
    else:
        group_rank = np.zeros(ng, dtype=np.int32)
    if not expand:
        anchor_rank = anchor_walk = anchor_vtx = np.zeros(0, dtype=np.int32)
        anchor_off = np.zeros(1, dtype=np.uint64)
    elif res.group_member_off:
        anchor_rank, anchor_walk, anchor_off, anchor_vtx = expand_groups(group_rank, group_len, group_vtx, member_off, member_walk)
    else:
        anchor_rank, anchor_walk, anchor_vtx = group_rank, member_walk, group_vtx
        anchor_off = np.concatenate([[0], np.cumsum(group_len, dtype=np.uint64)]).astype(np.uint64) if have and res.group_len else np.zeros(0, dtype=np.uint64)
    return IndexResultPy(
        count_sp_r=int(ns), n_walks=int(nw), n_filtered=int(res.n_filtered),
        spectrum=_np_from(res.spectrum, ns, np.uint64),
        anchor_rank=anchor_rank,
        anchor_walk=anchor_walk,
        anchor_off=anchor_off,
        anchor_vtx=anchor_vtx,
        minimizers_per_walk=_np_from(res.minimizers_per_walk, nw, np.uint64),
        anchors_per_walk=_np_from(res.anchors_per_walk, nw, np.uint64),
        read_kmer_positions=int(res.read_kmer_positions), path_kmer_positions=int(res.path_kmer_positions),
        read_minimizers_emitted=int(res.read_minimizers_emitted),
        path_minimizers_emitted=int(res.path_minimizers_emitted), path_hits=int(res.path_hits),
        n_walk_kmers=int(res.n_walk_kmers),
        shared_kmer_hist=_np_from(res.shared_kmer_hist, nw + 1, np.uint64) if res.shared_kmer_hist else None,
        n_groups=int(ng), group_rank=group_rank, group_len=group_len, group_vtx=group_vtx, group_member_off=member_off,
        member_walk=member_walk, member_walk_bytes=walk_bytes, rank_off=rank_off if len(rank_off) else None)
