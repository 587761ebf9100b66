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


class GraphView(C.Structure):
    _fields_ = [("n_vtx", C.c_uint32), ("seg_off", u64p), ("seg_bases", u8p), ("n_walks", C.c_uint32),
                ("walk_off", u64p), ("walk_vtx", u32p), ("top_order_map", i32p)]


class ReadsView(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("read_off", u64p), ("read_bases", u8p)]


class IndexParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("w", C.c_int32), ("threshold", C.c_float), ("debug", C.c_int32)]


class IndexResult(C.Structure):
    _fields_ = [("count_sp_r", C.c_int32), ("n_walks", C.c_uint32), ("n_filtered", C.c_int64),
                ("n_anchors", C.c_uint64), ("n_anchor_vtx", C.c_uint64),
                ("spectrum", u64p), ("rank_off", u64p), ("anchor_walk", i32p), ("anchor_len", u8p),
                ("anchor_vtx", i32p), ("minimizers_per_walk", u64p), ("anchors_per_walk", u64p),
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
    """Host copy of phi_index_result (numpy arrays).  anchor_rank / anchor_off are the per-anchor expansions of the ABI's
    compact rank_off / anchor_len arrays (what the reference-side adapter walks through)."""
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

    @property
    def n_anchors(self):
        return len(self.anchor_rank)

    def wire_bytes(self):
        """Bytes of the C result arrays (what crosses PCIe): spectrum, rank_off, anchor_walk, anchor_len, anchor_vtx, per-walk counters."""
        ns, na = self.count_sp_r, self.n_anchors
        return 8 * ns + 8 * (ns + 1) + 4 * na + na + 4 * len(self.anchor_vtx) + 16 * self.n_walks

    def anchors(self):
        """[(rank, walk, [vertices])] in final order."""
        off = self.anchor_off
        return [(int(self.anchor_rank[a]), int(self.anchor_walk[a]), self.anchor_vtx[off[a]:off[a + 1]].tolist())
                for a in range(self.n_anchors)]


def _np_from(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


def result_to_py(res: IndexResult) -> IndexResultPy:
    na, nv, nw, ns = res.n_anchors, res.n_anchor_vtx, res.n_walks, res.count_sp_r
    have = bool(res.anchor_len) or na == 0
    rank_off = _np_from(res.rank_off, ns + 1 if res.rank_off else 0, np.uint64)
    lens = _np_from(res.anchor_len, na, np.uint8)
    if len(rank_off):
        assert int(rank_off[-1]) == na and int(rank_off[0]) == 0
        anchor_rank = np.repeat(np.arange(ns, dtype=np.int32), np.diff(rank_off.astype(np.int64)))
    else:
        anchor_rank = np.zeros(na if have else 0, dtype=np.int32)
    anchor_off = np.concatenate([[0], np.cumsum(lens, dtype=np.uint64)]).astype(np.uint64) if res.anchor_len else np.zeros(0, dtype=np.uint64)
    return IndexResultPy(
        count_sp_r=int(ns), n_walks=int(nw), n_filtered=int(res.n_filtered),
        spectrum=_np_from(res.spectrum, ns, np.uint64),
        anchor_rank=anchor_rank,
        anchor_walk=_np_from(res.anchor_walk, na, np.int32),
        anchor_off=anchor_off,
        anchor_vtx=_np_from(res.anchor_vtx, nv, np.int32),
        minimizers_per_walk=_np_from(res.minimizers_per_walk, nw, np.uint64),
        anchors_per_walk=_np_from(res.anchors_per_walk, nw, np.uint64),
        read_kmer_positions=int(res.read_kmer_positions), path_kmer_positions=int(res.path_kmer_positions),
        read_minimizers_emitted=int(res.read_minimizers_emitted),
        path_minimizers_emitted=int(res.path_minimizers_emitted), path_hits=int(res.path_hits),
        n_walk_kmers=int(res.n_walk_kmers),
        shared_kmer_hist=_np_from(res.shared_kmer_hist, nw + 1, np.uint64) if res.shared_kmer_hist else None)
