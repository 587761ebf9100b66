"""ctypes binding of include/phi_gpu_index.h."""
import ctypes as C
import os

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class PhiGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"phi_gpu_index error {code}: {msg}")
        self.code = code


def library_path():
    return os.path.join(_HERE, "libphi_gpu_index.so")


def load_library():
    """Load libphi_gpu_index.so.  Raises if it has not been built (no fallback of any kind)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise PhiGpuError(-1, f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
    lib = C.CDLL(path)
    ctxp = C.c_void_p
    resp = C.POINTER(_abi.IndexResult)
    lib.phi_gpu_index_abi_version.restype = C.c_int
    lib.phi_gpu_index_create.restype = C.c_int
    lib.phi_gpu_index_create.argtypes = [C.c_int, C.POINTER(ctxp)]
    lib.phi_gpu_index_destroy.argtypes = [ctxp]
    lib.phi_gpu_last_error.restype = C.c_char_p
    lib.phi_gpu_last_error.argtypes = [ctxp]
    lib.phi_gpu_index_run.restype = C.c_int
    lib.phi_gpu_index_run.argtypes = [ctxp, C.POINTER(_abi.GraphView), C.POINTER(_abi.ReadsView),
                                      C.POINTER(_abi.IndexParams), C.POINTER(resp)]
    lib.phi_gpu_index_upload.restype = C.c_int
    lib.phi_gpu_index_upload.argtypes = [ctxp, C.POINTER(_abi.GraphView), C.POINTER(_abi.ReadsView)]
    lib.phi_gpu_index_run_resident.restype = C.c_int
    lib.phi_gpu_index_run_resident.argtypes = [ctxp, C.POINTER(_abi.IndexParams), C.c_int, C.POINTER(resp)]
    lib.phi_gpu_index_result_free.argtypes = [resp]
    lib.phi_gpu_index_last_times.restype = C.c_int
    lib.phi_gpu_index_last_times.argtypes = [ctxp, C.POINTER(_abi.StageTimes)]
    lib.phi_gpu_index_set_walk_sharing.restype = C.c_int
    lib.phi_gpu_index_set_walk_sharing.argtypes = [ctxp, C.c_int, C.c_int]
    lib.phi_gpu_index_last_sharing.restype = C.c_int
    lib.phi_gpu_index_last_sharing.argtypes = [ctxp, C.POINTER(_abi.SharingStats)]
    lib.phi_gpu_index_sketch_walks.restype = C.c_int
    lib.phi_gpu_index_sketch_walks.argtypes = [ctxp, C.POINTER(_abi.GraphView), C.POINTER(_abi.IndexParams),
                                               C.POINTER(resp), C.POINTER(_abi.u64p)]
    lib.phi_gpu_index_free_u64.argtypes = [_abi.u64p]
    lib.phi_gpu_hash128_to_64.restype = C.c_int
    lib.phi_gpu_hash128_to_64.argtypes = [ctxp, C.c_char_p, C.c_uint64, C.c_int32, _abi.u64p]
    lib.phi_gpu_host_alloc.restype = C.c_void_p
    lib.phi_gpu_host_alloc.argtypes = [C.c_size_t]
    lib.phi_gpu_host_free.argtypes = [C.c_void_p]
    lib.phi_gpu_index_comm_unique_id.restype = C.c_int
    lib.phi_gpu_index_comm_unique_id.argtypes = [C.c_char_p]
    lib.phi_gpu_index_comm_init.restype = C.c_int
    lib.phi_gpu_index_comm_init.argtypes = [ctxp, C.c_int, C.c_int, C.c_char_p, C.c_uint32, C.c_uint32]
    lib.phi_host_graph_load.restype = C.c_int
    lib.phi_host_graph_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    lib.phi_host_graph_view.restype = C.POINTER(_abi.GraphView)
    lib.phi_host_graph_view.argtypes = [C.c_void_p]
    lib.phi_host_graph_walk_name.restype = C.c_char_p
    lib.phi_host_graph_walk_name.argtypes = [C.c_void_p, C.c_uint32]
    lib.phi_host_graph_segment_name.restype = C.c_char_p
    lib.phi_host_graph_segment_name.argtypes = [C.c_void_p, C.c_uint32]
    lib.phi_host_graph_n_links.restype = C.c_uint64
    lib.phi_host_graph_n_links.argtypes = [C.c_void_p]
    lib.phi_host_graph_unlinked_steps.restype = C.c_uint64
    lib.phi_host_graph_unlinked_steps.argtypes = [C.c_void_p]
    lib.phi_host_graph_free.argtypes = [C.c_void_p]
    lib.phi_host_reads_load.restype = C.c_int
    lib.phi_host_reads_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    lib.phi_host_reads_view.restype = C.POINTER(_abi.ReadsView)
    lib.phi_host_reads_view.argtypes = [C.c_void_p]
    lib.phi_host_reads_name.restype = C.c_char_p
    lib.phi_host_reads_name.argtypes = [C.c_void_p, C.c_uint64]
    lib.phi_host_reads_free.argtypes = [C.c_void_p]
    lib.phi_shard_owner_of_hash.restype = C.c_int
    lib.phi_shard_owner_of_hash.argtypes = [C.c_uint64, C.c_int]
    lib.phi_shard_split_by_weight.restype = C.c_int
    lib.phi_shard_split_by_weight.argtypes = [_abi.u64p, C.c_uint64, C.c_int, _abi.u64p]
    lib.phi_shard_walk_regions.restype = C.c_int
    lib.phi_shard_walk_regions.argtypes = [C.POINTER(_abi.GraphView), C.c_int, _abi.u64p]
    lib.phi_shard_slice_walks.restype = C.c_int
    lib.phi_shard_slice_walks.argtypes = [C.POINTER(_abi.GraphView), C.c_int, C.c_int, C.c_uint64, C.c_uint64, _abi.u64p, _abi.u64p]
    lib.phi_shard_slice_walks_all.restype = C.c_int
    lib.phi_shard_slice_walks_all.argtypes = [C.POINTER(_abi.GraphView), C.c_int, C.c_int, C.c_int, _abi.u64p, _abi.u64p, _abi.u64p]
    lib.phi_gpu_index_set_walk_region.restype = C.c_int
    lib.phi_gpu_index_set_walk_region.argtypes = [ctxp, C.c_uint64, C.c_uint64]
    lib.phi_index_result_merge.restype = C.c_int
    lib.phi_index_result_merge.argtypes = [C.POINTER(resp), C.c_int, C.POINTER(resp)]
    _LIB = lib
    return lib


def load_gfa(path):
    """phi_host_graph_load: GFA (.gz or plain) -> Graph (numpy copies of the flat view) with walk and segment names."""
    lib = load_library()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib.phi_host_graph_load(os.fsencode(path), C.byref(h), err, len(err))
    if rc != _abi.PHI_OK:
        raise PhiGpuError(rc, err.value.decode())
    try:
        v = lib.phi_host_graph_view(h).contents
        nv, nw = v.n_vtx, v.n_walks
        seg_off = _abi._np_from(v.seg_off, nv + 1, np.uint64)
        walk_off = _abi._np_from(v.walk_off, nw + 1, np.uint64)
        g = _abi.Graph(seg_off, _abi._np_from(v.seg_bases, int(seg_off[-1]), np.uint8), walk_off,
                       _abi._np_from(v.walk_vtx, int(walk_off[-1]), np.uint32), _abi._np_from(v.top_order_map, nv, np.int32),
                       [lib.phi_host_graph_walk_name(h, i).decode() for i in range(nw)])
        g.segment_names = [lib.phi_host_graph_segment_name(h, i).decode() for i in range(nv)]
        g.n_links = int(lib.phi_host_graph_n_links(h))
        g.n_unlinked_steps = int(lib.phi_host_graph_unlinked_steps(h))
        return g
    finally:
        lib.phi_host_graph_free(h)


def load_reads(path):
    """phi_host_reads_load: FASTA / FASTQ (.gz or plain) -> (Reads, names)."""
    lib = load_library()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib.phi_host_reads_load(os.fsencode(path), C.byref(h), err, len(err))
    if rc != _abi.PHI_OK:
        raise PhiGpuError(rc, err.value.decode())
    try:
        v = lib.phi_host_reads_view(h).contents
        off = _abi._np_from(v.read_off, v.n_reads + 1, np.uint64)
        rd = _abi.Reads(off, _abi._np_from(v.read_bases, int(off[-1]), np.uint8))
        return rd, [lib.phi_host_reads_name(h, i).decode() for i in range(v.n_reads)]
    finally:
        lib.phi_host_reads_free(h)


class PhiGpuIndex:
    """One ctx == one GPU.  Mirrors the C ABI one to one."""

    def __init__(self, device=-1):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        rc = self.lib.phi_gpu_index_create(device, C.byref(self.ctx))
        if rc != _abi.PHI_OK:
            raise PhiGpuError(rc, self.lib.phi_gpu_last_error(None).decode())
        self._keep = None

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.phi_gpu_index_destroy(self.ctx)
            self.ctx = None
            for p in getattr(self, "_pinned", []):
                self.lib.phi_gpu_host_free(p)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _abi.PHI_OK:
            raise PhiGpuError(rc, self.lib.phi_gpu_last_error(self.ctx).decode())

    @staticmethod
    def _params(k, w, threshold, debug=0):
        return _abi.IndexParams(int(k), int(w), float(threshold), int(debug))

    def _take(self, resp, free=True, expand=True):
        res = _abi.result_to_py(resp.contents, expand)
        if free:
            self.lib.phi_gpu_index_result_free(resp)
        return res

    def run(self, graph, reads, k=31, w=25, threshold=1.0, debug=0, expand=True):
        """Host buffers in, host result out (H2D + kernels + D2H): the drop-in call.  Returns numpy copies."""
        return self._take(self.run_raw(graph, reads, k, w, threshold, debug), expand=expand)

    def run_raw(self, graph, reads, k=31, w=25, threshold=1.0, debug=0):
        """The same call, returning the C result pointer untouched (no numpy copies); free it with free_raw()."""
        gv, rv, prm = graph.view(), reads.view(), self._params(k, w, threshold, debug)
        out = C.POINTER(_abi.IndexResult)()
        self._check(self.lib.phi_gpu_index_run(self.ctx, C.byref(gv), C.byref(rv), C.byref(prm), C.byref(out)))
        return out

    def free_raw(self, out):
        self.lib.phi_gpu_index_result_free(out)

    def pinned_copy(self, arr):
        """Copy a numpy array into pinned host memory (phi_gpu_host_alloc) and return a numpy view on it."""
        arr = np.ascontiguousarray(arr)
        nbytes = max(arr.nbytes, 1)
        p = self.lib.phi_gpu_host_alloc(nbytes)
        if not p:
            raise PhiGpuError(_abi.PHI_ERR_NOMEM, "phi_gpu_host_alloc failed")
        buf = (C.c_uint8 * nbytes).from_address(p)
        out = np.frombuffer(buf, dtype=arr.dtype, count=arr.size).reshape(arr.shape)
        out[...] = arr
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        return out

    def pinned_inputs(self, graph, reads):
        """Graph / Reads whose arrays live in pinned host memory."""
        g = _abi.Graph(self.pinned_copy(graph.seg_off), self.pinned_copy(graph.seg_bases), self.pinned_copy(graph.walk_off),
                       self.pinned_copy(graph.walk_vtx), self.pinned_copy(graph.top_order_map), graph.walk_names)
        r = _abi.Reads(self.pinned_copy(reads.read_off), self.pinned_copy(reads.read_bases))
        return g, r

    def upload(self, graph, reads):
        gv, rv = graph.view(), reads.view()
        self._check(self.lib.phi_gpu_index_upload(self.ctx, C.byref(gv), C.byref(rv)))

    def run_resident(self, k=31, w=25, threshold=1.0, download=True):
        prm = self._params(k, w, threshold)
        out = C.POINTER(_abi.IndexResult)()
        self._check(self.lib.phi_gpu_index_run_resident(self.ctx, C.byref(prm), 1 if download else 0, C.byref(out)))
        return self._take(out)

    def times(self):
        t = _abi.StageTimes()
        self._check(self.lib.phi_gpu_index_last_times(self.ctx, C.byref(t)))
        return t.as_dict()

    def set_walk_sharing(self, chunk_shift=11, share=True):
        self._check(self.lib.phi_gpu_index_set_walk_sharing(self.ctx, int(chunk_shift), 1 if share else 0))

    def set_walk_region(self, coord_lo, coord_hi):
        """Own only the windows whose last k-mer starts on a vertex with topological base coordinate in [coord_lo, coord_hi)."""
        self._check(self.lib.phi_gpu_index_set_walk_region(self.ctx, int(coord_lo), int(coord_hi)))

    def sharing(self):
        t = _abi.SharingStats()
        self._check(self.lib.phi_gpu_index_last_sharing(self.ctx, C.byref(t)))
        return t.as_dict()

    def sketch_walks(self, graph, k=31, w=25):
        """ILP_index::index_kmers for every walk: (result in (walk, path) order, hashes)."""
        gv, prm = graph.view(), self._params(k, w, 1.0)
        out = C.POINTER(_abi.IndexResult)()
        hp = _abi.u64p()
        self._check(self.lib.phi_gpu_index_sketch_walks(self.ctx, C.byref(gv), C.byref(prm), C.byref(out), C.byref(hp)))
        res = _abi.result_to_py(out.contents)
        hashes = _abi._np_from(hp, res.n_anchors, np.uint64)
        self.lib.phi_gpu_index_result_free(out)
        self.lib.phi_gpu_index_free_u64(hp)
        return res, hashes

    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id (rank 0 creates it, the caller distributes it)."""
        lib = load_library()
        buf = C.create_string_buffer(_abi.PHI_COMM_ID_BYTES)
        rc = lib.phi_gpu_index_comm_unique_id(buf)
        if rc != _abi.PHI_OK:
            raise PhiGpuError(rc, lib.phi_gpu_last_error(None).decode())
        return buf.raw

    def comm_init(self, rank, world, unique_id, walk_id_base, n_walks_global):
        self._check(self.lib.phi_gpu_index_comm_init(self.ctx, rank, world, unique_id, walk_id_base, n_walks_global))

    def hash128_to_64(self, keys, length):
        """Device MurmurHash3_x64_128 -> h0^h1 over len(keys)//length packed keys."""
        n = len(keys) // length
        out = np.zeros(n, dtype=np.uint64)
        self._check(self.lib.phi_gpu_hash128_to_64(self.ctx, bytes(keys), n, length, out.ctypes.data_as(_abi.u64p)))
        return out
