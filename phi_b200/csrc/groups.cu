// Grouped result of the threshold filter (hand-written sm_100a CUDA).
//
// The reference's filter (/root/reference/src/ILP_index.cpp:670-716) builds, per rank, a std::map from the key string of a
// vertex list to (count, [(walk, anchor) ...]) and then re-emits, group after group in key order, one anchor per member into
// Anchor_hits[rank][walk].  That map IS the natural result: one vertex list per group plus the walks that carry it.  The
// device keeps it in that form instead of instantiating one record per (walk, anchor): with h haplotypes the per-anchor form
// repeats every vertex list up to h times, and it is the form that has to cross PCIe.
//
//   group table (filter.cu)            : slot = (rank, vertex list), count, and for every hit of a representative chunk the
//                                        count before it was added (its sub-offset inside the group)
//   surviving groups -> (rank, hit id) : compaction, stable radix sort on the rank, key order inside a rank (filter.cu)
//   sizes -> scans                     : member and vertex offsets of every group, first group of every rank
//   fill                               : every surviving hit copies the member walks of its chunk (chunks.cu keeps them,
//                                        ascending, per representative) to its sub-offset; the group's first hit writes the list
//   members                            : one warp per group merges the parts (counting sort on the walk id), adds the walk id
//                                        base of this GPU and counts the anchors per walk
#include "kernels.h"
#include "device_common.cuh"
#include <algorithm>

namespace phi {

#define PHI_LAUNCH_CHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return e_; if (launches) ++*launches; } while (0)

constexpr uint32_t C_NONE = 0xFFFFFFFFu;
constexpr int SORT_WARPS = 4;                    // warps per block of the sorted-copy kernels
constexpr uint32_t G_DROPPED = 0xFFFFFFFFu;      // g_rep of a group whose rank was dropped (set once the groups are compacted)
constexpr uint32_t SUB_REP_BIT = 0x80000000u;    // hit_sub: this hit represents its group (writes the vertex list)
constexpr int GROUPS_PER_WARP = 16;              // group_members_kernel: groups per warp (one histogram flush per block)

// ---- src[0, n) -> dst[0, n) ascending (+ add), values < n_vals, one warp.  sorted: plain copy.  Otherwise a counting sort over
// the warp's cnt[min(n_vals, SORT_VALS_MAX)] shared counters, one pass per range of SORT_VALS_MAX values.
template <class OutT, class Hist>
__device__ __forceinline__ void warp_sorted_copy(const uint32_t *src, OutT *dst, uint32_t n, bool sorted, uint32_t n_vals, uint32_t add,
                                                 uint32_t *cnt, int lane, Hist hist)
{
    if (sorted || n < 2) {
        for (uint32_t i = lane; i < n; i += 32) { const uint32_t v = src[i] + add; dst[i] = (OutT)v; hist(v, 1u); }
        return;
    }
    uint32_t base = 0;
    for (uint32_t r0 = 0; r0 < n_vals && base < n; r0 += SORT_VALS_MAX) {
        const uint32_t span = min(n_vals - r0, SORT_VALS_MAX);
        for (uint32_t v = lane; v < span; v += 32) cnt[v] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) { const uint32_t v = src[i] - r0; if (v < span) atomicAdd(&cnt[v], 1u); }
        __syncwarp();
        for (uint32_t v0 = 0; v0 < span && base < n; v0 += 32) {
            const uint32_t v = v0 + lane, c = v < span ? cnt[v] : 0u;
            uint32_t inc = c;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
            for (uint32_t q = 0; q < c; ++q) dst[base + inc - c + q] = (OutT)(r0 + v + add);
            if (c) hist(r0 + v + add, c);
            base += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        __syncwarp();
    }
}

struct NoHist { __device__ __forceinline__ void operator()(uint32_t, uint32_t) const {} };

// ------------------------------------------------------------------ member walks of every representative chunk
__global__ void chunk_member_fill_kernel(ChunkTable C, const uint32_t *cm_off, uint32_t *cursor, uint32_t *cm_tmp)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C.n_chunks) return;
    const uint32_t rep = C.c_rep[c];
    if (rep == C_NONE) return;
    cm_tmp[cm_off[rep] + atomicAdd(&cursor[rep], 1u)] = C.c_walk[c];
}
__global__ void __launch_bounds__(SORT_WARPS * 32) chunk_member_sort_kernel(ChunkTable C, const uint32_t *cm_off, const uint32_t *cm_tmp, uint32_t *cm_walk,
                                                                          uint32_t n_vals)
{
    extern __shared__ uint32_t s_cnt[];
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (c >= C.n_chunks) return;
    const uint32_t off = cm_off[c], n = cm_off[c + 1] - off;
    if (!n) return;
    warp_sorted_copy(cm_tmp + off, cm_walk + off, n, false, n_vals, 0u, s_cnt + (size_t)wid * min(n_vals, SORT_VALS_MAX), lane, NoHist());
}

cudaError_t chunk_members(const ChunkTable &C, uint32_t n_walks, uint32_t *cm_off, uint32_t *cursor, uint32_t *cm_tmp, uint32_t *cm_walk,
                          void *scan_scratch, cudaStream_t st, uint64_t *launches)
{
    if (!C.n_chunks) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(C.c_ninst + C.n_chunks, 0, 4, st);
    if (e != cudaSuccess) return e;
    e = scan_u32(C.c_ninst, cm_off, (uint64_t)C.n_chunks + 1, scan_scratch, st, launches);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(cursor, 0, (size_t)C.n_chunks * 4, st);
    if (e != cudaSuccess) return e;
    chunk_member_fill_kernel<<<(C.n_chunks + 255) / 256, 256, 0, st>>>(C, cm_off, cursor, cm_tmp);
    PHI_LAUNCH_CHECK();
    chunk_member_sort_kernel<<<(unsigned)(((uint64_t)C.n_chunks + SORT_WARPS - 1) / SORT_WARPS), SORT_WARPS * 32,
                               (size_t)SORT_WARPS * std::min(n_walks, SORT_VALS_MAX) * 4, st>>>(C, cm_off, cm_tmp, cm_walk, n_walks);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

// ------------------------------------------------------------------ surviving groups
// A group is represented by the hit that claimed its slot.  flags[i] = 1 for the representing hit of a group whose rank survives;
// totals of groups / members / vertices go to the counter block.
__global__ void __launch_bounds__(256) group_flags_kernel(FilterArgs A, FilterWork W, uint32_t *flags)
{
    __shared__ unsigned long long s_tot[3];
    if (threadIdx.x < 3) s_tot[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t f = 0; unsigned long long members = 0, vtx = 0;
    if (i < A.n_hits) {
        // most hits belong to dropped ranks (minimizers every walk carries): the drop flags (one byte per rank, L2-resident) are looked
        // at first, the scattered slot of the group table only for the hits that survive
        if (!W.rank_drop[A.hit_rank[i]]) {
            const uint2 gs = W.g_slot[W.hit_slot[i]];
            if (gs.x == (uint32_t)i) { f = 1; members = gs.y; vtx = A.hit_nv[i]; }
        }
        flags[i] = f;
    }
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, f != 0);
    if (b) {
        #pragma unroll
        for (int d = 16; d; d >>= 1) { members += __shfl_xor_sync(0xFFFFFFFFu, members, d); vtx += __shfl_xor_sync(0xFFFFFFFFu, vtx, d); }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&s_tot[0], (unsigned long long)__popc(b)); atomicAdd(&s_tot[1], members); atomicAdd(&s_tot[2], vtx); }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_tot[0]) {
        atomicAdd(&W.ctr[CTR_OUT_GROUPS], s_tot[0]); atomicAdd(&W.ctr[CTR_SURVIVORS], s_tot[1]); atomicAdd(&W.ctr[CTR_SURV_VTX], s_tot[2]);
    }
}
__global__ void group_emit_kernel(FilterArgs A, const uint32_t *flags, const uint32_t *pos, uint32_t *keys, uint32_t *vals)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < A.n_hits && flags[i]) { keys[pos[i]] = A.hit_rank[i]; vals[pos[i]] = (uint32_t)i; }
}

cudaError_t groups_compact(const FilterArgs &A, const FilterWork &W, uint32_t *flags, uint32_t *pos, uint32_t *keys, uint32_t *vals,
                           void *scan_scratch, cudaStream_t st, uint64_t *launches)
{
    if (!A.n_hits) return cudaSuccess;
    const unsigned nb = (unsigned)((A.n_hits + 255) / 256);
    group_flags_kernel<<<nb, 256, 0, st>>>(A, W, flags);
    PHI_LAUNCH_CHECK();
    cudaError_t e = scan_u32(flags, pos, A.n_hits, scan_scratch, st, launches);
    if (e != cudaSuccess) return e;
    group_emit_kernel<<<nb, 256, 0, st>>>(A, flags, pos, keys, vals);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

// per group, in final order: members, vertices, the slot's output index (g_rep is free by now), first group of every rank
__global__ void group_sizes_kernel(FilterArgs A, FilterWork W, const uint32_t *order, uint32_t n, uint32_t *cnt_out, uint32_t *nv_out,
                                   uint8_t *group_len, uint32_t *rank_off)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t i = order[j], slot = W.hit_slot[i];
    const uint8_t nv = A.hit_nv[i];
    cnt_out[j] = W.g_slot[slot].y; nv_out[j] = nv; group_len[j] = nv;
    W.g_slot[slot].x = j;
    W.hit_sub[i] |= SUB_REP_BIT;
    const int64_t r = A.hit_rank[i], rp = j ? (int64_t)A.hit_rank[order[j - 1]] : -1;
    for (int64_t q = rp + 1; q <= r; ++q) rank_off[q] = j;
    if (j == n - 1) for (int64_t q = r + 1; q <= (int64_t)A.n_ranks; ++q) rank_off[q] = n;
}

// 4 lanes per surviving hit: the member walks of its chunk go to the hit's sub-offset inside its group; the representing hit
// also writes the vertex list
__global__ void __launch_bounds__(256) group_fill_kernel(FilterArgs A, FilterWork W, GroupOut G)
{
    const uint64_t i = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 2;
    const uint32_t sub = threadIdx.x & 3;
    if (i >= A.n_hits) return;
    if (W.rank_drop[A.hit_rank[i]]) return;                               // (before the scattered read below: most hits stop here)
    const uint32_t j = W.g_slot[W.hit_slot[i]].x;                         // output index of the hit's group
    const uint32_t c = A.hit_walk[i];                                     // hits of representatives: the chunk takes the place of the walk
    const uint32_t s0 = G.cm_off[c], n = G.cm_off[c + 1] - s0;
    const uint32_t hs = W.hit_sub[i];
    const uint32_t *src = G.cm_walk + s0;
    uint32_t *dst = G.members_tmp + G.member_off[j] + (hs & ~SUB_REP_BIT);
    for (uint32_t q = sub; q < n; q += 4) dst[q] = src[q];
    if (hs & SUB_REP_BIT) {
        const int32_t *p = A.vtx_pool + A.hit_voff[i];
        int32_t *o = G.group_vtx + G.vtx_off[j];
        const uint32_t nv = A.hit_nv[i];
        for (uint32_t q = sub; q < nv; q += 4) o[q] = p[q];
    }
}

// one warp per group: parts -> ascending member walks (global walk ids), anchors per walk
template <class OutT>
__global__ void __launch_bounds__(SORT_WARPS * 32) group_members_kernel(FilterArgs A, FilterWork W, GroupOut G, uint32_t n_groups, uint32_t n_vals,
                                                                      uint32_t n_walks_out, int use_hist)
{
    extern __shared__ uint32_t s_mem[];
    uint32_t *s_hist = s_mem;                                             // [n_walks_out] when use_hist
    uint32_t *s_cnt = s_mem + (use_hist ? n_walks_out : 0);               // [SORT_WARPS][min(n_vals, SORT_VALS_MAX)]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (use_hist) { for (uint32_t q = threadIdx.x; q < n_walks_out; q += blockDim.x) s_hist[q] = 0; __syncthreads(); }
    unsigned long long *g_hist = G.anchors_per_walk;
    auto hist = [&](uint32_t v, uint32_t cnt) { if (use_hist) atomicAdd(&s_hist[v], cnt); else atomicAdd(&g_hist[v], (unsigned long long)cnt); };
    const uint32_t j0 = (blockIdx.x * SORT_WARPS + wid) * GROUPS_PER_WARP;
    for (uint32_t j = j0; j < j0 + GROUPS_PER_WARP && j < n_groups; ++j) {
        const uint32_t i = G.order[j], off = G.member_off[j], n = G.member_off[j + 1] - off;
        const uint32_t c = A.hit_walk[i];
        const bool single = G.cm_off[c + 1] - G.cm_off[c] == n;            // one part: the chunk's member list, already ascending
        warp_sorted_copy(G.members_tmp + off, (OutT *)G.member_walk + off, n, single, n_vals, G.walk_id_base,
                         s_cnt + (size_t)wid * min(n_vals, SORT_VALS_MAX), lane, hist);
    }
    if (use_hist) {
        __syncthreads();
        for (uint32_t q = threadIdx.x; q < n_walks_out; q += blockDim.x) if (s_hist[q]) atomicAdd(&G.anchors_per_walk[q], (unsigned long long)s_hist[q]);
    }
}

cudaError_t groups_sizes(const FilterArgs &A, const FilterWork &W, const uint32_t *order, uint32_t n_groups, uint32_t *cnt_out, uint32_t *nv_out,
                         uint8_t *group_len, uint32_t *rank_off, cudaStream_t st, uint64_t *launches)
{
    if (!n_groups) return cudaSuccess;
    group_sizes_kernel<<<(n_groups + 255) / 256, 256, 0, st>>>(A, W, order, n_groups, cnt_out, nv_out, group_len, rank_off);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t groups_fill(const FilterArgs &A, const FilterWork &W, const GroupOut &G, uint32_t n_groups, uint32_t n_walks_local, uint32_t n_walks_out,
                        cudaStream_t st, uint64_t *launches)
{
    if (!n_groups || !A.n_hits) return cudaSuccess;
    group_fill_kernel<<<(unsigned)((A.n_hits * 4 + 255) / 256), 256, 0, st>>>(A, W, G);
    PHI_LAUNCH_CHECK();
    const int use_hist = n_walks_out <= GROUP_HIST_MAX;
    const size_t smem = ((use_hist ? (size_t)n_walks_out : 0) + (size_t)SORT_WARPS * std::min(n_walks_local, SORT_VALS_MAX)) * 4;
    const unsigned nb = (n_groups + SORT_WARPS * GROUPS_PER_WARP - 1) / (SORT_WARPS * GROUPS_PER_WARP);
    if (G.member_walk_bytes == 2) group_members_kernel<uint16_t><<<nb, SORT_WARPS * 32, smem, st>>>(A, W, G, n_groups, n_walks_local, n_walks_out, use_hist);
    else group_members_kernel<uint32_t><<<nb, SORT_WARPS * 32, smem, st>>>(A, W, G, n_groups, n_walks_local, n_walks_out, use_hist);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

}  // namespace phi
