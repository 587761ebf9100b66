"""Multi-GPU host logic: sharding of reads / walks over ranks and merge of the per-rank results (thin wrappers over the
library's phi_shard_* / phi_index_result_merge).  torch.distributed is plumbing only (unique-id hand-round, barrier,
max-over-ranks of the timings, in bench.py and the tests); the data exchange itself is NCCL inside libphi_gpu_index.so."""
import ctypes as C

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


def region_bounds(graph, world):
    """phi_shard_walk_regions: world+1 bounds of the topological base coordinate, balanced by the steps of the given walks.
    Returns None when the graph does not allow a region cut (top_order_map is not a permutation)."""
    lib = load_library()
    b = np.zeros(world + 1, dtype=np.uint64)
    gv = graph.view()
    rc = lib.phi_shard_walk_regions(C.byref(gv), world, b.ctypes.data_as(_abi.u64p))
    if rc == _abi.PHI_ERR_UNSUPPORTED:
        return None
    assert rc == 0, rc
    return b


def slice_walks(graph, k, w, coord_lo, coord_hi):
    """phi_shard_slice_walks: the graph view of the GPU that owns [coord_lo, coord_hi): every walk cut to its steps inside the range
    plus context.  Returns None when a walk does not follow the topological order."""
    lib = load_library()
    first = np.zeros(graph.n_walks, dtype=np.uint64)
    length = np.zeros(graph.n_walks, dtype=np.uint64)
    gv = graph.view()
    rc = lib.phi_shard_slice_walks(C.byref(gv), k, w, int(coord_lo), int(coord_hi), first.ctypes.data_as(_abi.u64p), length.ctypes.data_as(_abi.u64p))
    if rc == _abi.PHI_ERR_UNSUPPORTED:
        return None
    assert rc == 0, rc
    from .synth import expand_ranges
    vtx = graph.walk_vtx[expand_ranges(first.astype(np.int64), length.astype(np.int64))]
    off = np.concatenate([[0], np.cumsum(length.astype(np.int64))]).astype(np.uint64)
    return _abi.Graph(graph.seg_off, graph.seg_bases, off, vtx, graph.top_order_map, graph.walk_names)


def slice_walks_all(graph, k, w, bounds):
    """phi_shard_slice_walks_all: (slice_first, slice_len), each [world][n_walks], for all regions of `bounds` at once (the order
    check of the walks runs once).  None when a walk does not follow the topological order."""
    lib = load_library()
    world = len(bounds) - 1
    first = np.zeros((world, graph.n_walks), dtype=np.uint64)
    length = np.zeros((world, graph.n_walks), dtype=np.uint64)
    b = np.ascontiguousarray(bounds, dtype=np.uint64)
    gv = graph.view()
    rc = lib.phi_shard_slice_walks_all(C.byref(gv), k, w, world, b.ctypes.data_as(_abi.u64p), first.ctypes.data_as(_abi.u64p), length.ctypes.data_as(_abi.u64p))
    if rc == _abi.PHI_ERR_UNSUPPORTED:
        return None
    assert rc == 0, rc
    return first, length


def shard_inputs(graph, reads, rank, world, k=31, w=25, mode="region"):
    """This rank's shard.  Reads: contiguous, balanced by bases.  Walks, mode "region" (default): ALL walks cut to this rank's range
    of the topological base coordinate (walk sharing keeps working: identical chunks of different walks meet on one GPU);
    mode "walk", or a graph that does not allow a region cut: contiguous whole walks balanced by steps.  Segments and
    top_order_map are replicated.  Returns (graph_shard, reads_shard, walk_id_base, region) with region = (coord_lo, coord_hi) or None."""
    rb = split_by_weight(reads.read_off, world)
    rs = reads.take(int(rb[rank]), int(rb[rank + 1]))
    if mode == "region":
        b = region_bounds(graph, world)
        if b is not None:
            gs = slice_walks(graph, k, w, b[rank], b[rank + 1])
            if gs is not None:
                return gs, rs, 0, (int(b[rank]), int(b[rank + 1]))
    wb = split_by_weight(graph.walk_off, world)
    return graph.take_walks(int(wb[rank]), int(wb[rank + 1])), rs, int(wb[rank]), None


def merge_results(parts, expand=True):
    """Per-rank results -> the global result (phi_index_result_merge): per hash rank the groups of all parts in the reference's key
    order, the member lists of a group that several GPUs hold united; per-walk counters, n_filtered and work counters summed; the
    spectrum from the part that carries it."""
    lib = load_library()
    cs = [_abi.py_to_c_result(p) for p in parts]
    arr = (C.POINTER(_abi.IndexResult) * len(parts))(*[C.pointer(c[0]) for c in cs])
    out = C.POINTER(_abi.IndexResult)()
    import time
    t0 = time.perf_counter()
    rc = lib.phi_index_result_merge(arr, len(parts), C.byref(out))
    merge_results.last_call_ms = (time.perf_counter() - t0) * 1e3       # the library call alone (the numpy conversions around it are test / bench plumbing)
    assert rc == 0, rc
    res = _abi.result_to_py(out.contents, expand)
    lib.phi_gpu_index_result_free(out)
    return res


def init_comm(ix, rank, world, walk_id_base, n_walks_global, dist):
    """Create the library's NCCL communicator: rank 0 makes the id, torch.distributed hands it round."""
    box = [ix.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ix.comm_init(rank, world, box[0], walk_id_base, n_walks_global)
