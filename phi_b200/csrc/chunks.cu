// Walk chunking and de-duplication (hand-written sm_100a CUDA).
//
// The reference sketches every haplotype walk independently (/root/reference/src/ILP_index.cpp:556-573,
// one index_kmers(h) per walk), although the walks of a pangenome graph share most of their vertex
// sequences.  What index_kmers emits for the windows ending inside a stretch of a walk is a pure function
// of the vertex sequence under that stretch plus w bases of left and k-1 bases of right context
// (:388-442: the k-mers of the windows, the previous window's minimum for the prev_hash chain, and the
// vertices under the chosen k-mer).  So walks are cut into CHUNKS at content-defined boundaries (a step
// starts a chunk when its vertex falls into a new bucket of the topological base coordinate — the same
// vertex starts a chunk in every walk that reaches it the same way), chunks with identical context are
// grouped, ONE representative per group is sketched and probed, and its hits are instantiated for every
// member with the member's walk id and base offset.  Bit-exactness does not depend on where the
// boundaries fall or on how much sharing there is; only the amount of work does.
//
// Pipeline (all on the ctx stream):
//   topo coordinate -> boundary flags -> chunk table -> geometry + 128-bit fingerprint (warp per chunk)
//   -> open-addressing grouping (smallest chunk id represents) -> exact verification of every member
//   against its representative -> tile records of the representatives.
#include "kernels.h"
#include "device_common.cuh"
#include <cstdlib>

namespace phi {

#define PHI_LAUNCH_CHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return e_; if (launches) ++*launches; } while (0)

constexpr uint32_t C_NONE = 0xFFFFFFFFu;

__device__ __forceinline__ uint64_t cmix(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// ---- topological base coordinate of every vertex: bases of all vertices that precede it in top_order_map
__global__ void topo_len_kernel(const int32_t *top_order_map, const uint64_t *seg_off, uint32_t n_vtx, uint32_t *tlen, unsigned long long *ctr)
{
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vtx) return;
    int32_t t = top_order_map[v];
    if (t < 0 || (uint32_t)t >= n_vtx) { ctr[CTR_BAD_TOPO] = 1; return; }
    tlen[t] = (uint32_t)(seg_off[v + 1] - seg_off[v]);
}
// what the step pass needs of a vertex, in one 16-byte record: (bases, coordinate bucket, top_order_map, -).
// coordinate = prefix[top_order_map[v]]; with an unusable top_order_map the segment-store offset serves (any function of v is correct)
__global__ void topo_coord_kernel(const int32_t *top_order_map, const uint64_t *seg_off, const uint64_t *prefix, uint32_t n_vtx,
                                  int shift, uint64_t own_lo, uint64_t own_hi, unsigned long long *ctr, uint4 *vinfo)
{
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vtx) return;
    const uint64_t len = seg_off[v + 1] - seg_off[v];
    if (len >= (1ull << 31)) ctr[CTR_SEG_TOO_LONG] = 1;
    const uint64_t coord = ctr[CTR_BAD_TOPO] ? seg_off[v] : prefix[top_order_map[v]];
    const uint32_t region = (coord >= own_lo ? 1u : 0u) + (coord >= own_hi ? 1u : 0u);   // 1: owned by this GPU
    vinfo[v] = make_uint4((uint32_t)len, (uint32_t)(coord >> shift), (uint32_t)top_order_map[v], region);
}

__device__ __forceinline__ uint32_t walk_of_step(const uint64_t *walk_off, uint32_t n_walks, uint64_t s)
{
    uint32_t lo = 0, hi = n_walks;                                   // last h with walk_off[h] <= s
    while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (walk_off[m] <= s) lo = m; else hi = m; }
    return lo;
}

// last h in [0, n_walks) with walk_off[h] <= s, by one warp: 32 evenly spaced probes per round
__device__ __forceinline__ uint32_t walk_of_step_warp(const uint64_t *walk_off, uint32_t n_walks, uint64_t s, int lane)
{
    uint32_t lo = 0, hi = n_walks;                                   // walk_off[lo] <= s, walk_off[hi] > s (hi == n_walks: virtual)
    while (hi - lo > 1) {
        const uint32_t step = (hi - lo + 31) / 32;
        const uint64_t idx = (uint64_t)lo + (uint64_t)(lane + 1) * step;
        const bool le = idx < hi && walk_off[idx] <= s;              // monotone over the lanes
        const uint32_t c = __popc(__ballot_sync(0xFFFFFFFFu, le));
        lo += c * step;
        hi = min(hi, lo + step);
    }
    return lo;
}
// The steps [s0, s1] of one block mostly lie in one walk: warps 0 and 1 look the two ends up, the block shares the answer and
// only blocks that straddle a walk boundary search per thread (between the two ends).  Returns the walk of step s; ws = its first step.
__device__ __forceinline__ uint32_t walk_of_step_block(const uint64_t *walk_off, uint32_t n_walks, uint64_t s, uint64_t s0, uint64_t s1, uint64_t &ws)
{
    __shared__ uint32_t sh_h[2];
    __shared__ uint64_t sh_ws;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (wid < 2) {
        const uint32_t a = walk_of_step_warp(walk_off, n_walks, wid ? s1 : s0, lane);
        if (lane == 0) { sh_h[wid] = a; if (!wid) sh_ws = walk_off[a]; }
    }
    __syncthreads();
    uint32_t lo = sh_h[0], hi = sh_h[1];
    if (lo == hi) { ws = sh_ws; return lo; }
    ++hi;                                                               // last h in [lo, hi) with walk_off[h] <= s
    while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (walk_off[m] <= s) lo = m; else hi = m; }
    ws = walk_off[lo];
    return lo;
}

// ---- one pass over the walk steps: segment length, chunk boundary (the step's vertex lies in another coordinate bucket than the
// previous step's, or the step is the first of its walk), zero-length steps, topological monotonicity of the walks
// vertex record of a step; an id outside [0, n_vtx) (caller error) is flagged and reads as an empty vertex
__device__ __forceinline__ uint4 vinfo_of(const uint4 *vinfo, uint32_t n_vtx, uint32_t v, unsigned long long *ctr)
{
    if (v < n_vtx) return vinfo[v];
    ctr[CTR_BAD_VTX] = 1;
    return make_uint4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(256) step_pass_kernel(const uint32_t *walk_vtx, const uint64_t *walk_off, uint32_t n_walks, uint64_t n_steps,
                                                        const uint4 *vinfo, uint32_t n_vtx, PackedStep *packed, unsigned long long *ctr)
{
    const uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool zero = false, flag = false;
    uint4 me = make_uint4(0, 0, 0, 0);
    if (s < n_steps) me = vinfo_of(vinfo, n_vtx, walk_vtx[s], ctr);   // one 16-byte gather per step
    // the previous step's vertex record: the neighbouring lane has it (lane 0 fetches its own)
    uint32_t pb = __shfl_up_sync(0xFFFFFFFFu, me.y, 1), pt = __shfl_up_sync(0xFFFFFFFFu, me.z, 1), pr = __shfl_up_sync(0xFFFFFFFFu, me.w, 1);
    if (lane == 0 && s && s < n_steps) { const uint4 p = vinfo_of(vinfo, n_vtx, walk_vtx[s - 1], ctr); pb = p.y; pt = p.z; pr = p.w; }
    uint64_t ws;
    const uint64_t blk0 = blockIdx.x * (uint64_t)blockDim.x;
    walk_of_step_block(walk_off, n_walks, s < n_steps ? s : n_steps - 1, blk0, min(blk0 + blockDim.x, n_steps) - 1, ws);
    if (s < n_steps) {
        flag = s == ws;
        if (!flag) {
            flag = me.y != pb || me.w != pr;                          // another coordinate bucket, or across the owned range's border
            if ((int32_t)pt >= (int32_t)me.z) ctr[CTR_NONMONO] = 1;
        }
        zero = me.x == 0;
        PackedStep p; p.v = (me.x & 0x7FFFFFFFu) | (flag ? 0x80000000u : 0u);
        packed[s] = p;
    }
    const uint32_t bz = __ballot_sync(0xFFFFFFFFu, zero), bf = __ballot_sync(0xFFFFFFFFu, flag);
    if (lane == 0) {
        if (bz) atomicAdd(&ctr[CTR_ZERO_STEPS], (unsigned long long)__popc(bz));
        if (bf) atomicAdd(&ctr[CTR_CHUNK_FLAGS], (unsigned long long)__popc(bf));
    }
}

// ---- the same pass, the scan and the finalisation in ONE kernel (the common case: no zero-length steps).  Every tile of
// FS_TILE steps scans (chunk flag, bases since the walk start) in registers and gets the total of the tiles before it by a
// decoupled look-back over one 64-bit word per tile; nothing but walk_vtx is read and nothing but step_base (and the few
// per-chunk / per-walk values) is written.  The scanned value: bits [0,35) bases since the last walk start, [35,61) chunk
// starts, bit 61 "a walk started" (the bases of the left operand are discarded); bits 62-63 of a tile word: 1 = aggregate, 2 = prefix.
// every thread owns FS_ITEMS consecutive steps (4, 8 or 16; PHI_GPU_FS_ITEMS), a tile has FS_THREADS threads (128 or 256; PHI_GPU_FS_THREADS)
constexpr uint64_t FS_BASES = (1ull << 35) - 1, FS_CHUNKS = ((1ull << 26) - 1) << 35, FS_RESET = 1ull << 61, FS_VALUE = (1ull << 62) - 1;
__device__ __forceinline__ uint64_t fs_comb(uint64_t a, uint64_t b)           // a, then b
{
    const uint64_t chunks = ((a & FS_CHUNKS) + (b & FS_CHUNKS)) & FS_CHUNKS;
    if (b & FS_RESET) return (b & FS_BASES) | chunks | FS_RESET;
    return (((a & FS_BASES) + (b & FS_BASES)) & FS_BASES) | chunks | (a & FS_RESET);
}
template <int FS_ITEMS, int FS_THREADS>
__global__ void __launch_bounds__(FS_THREADS, (FS_ITEMS >= 16 ? 2 : 4) * (256 / FS_THREADS)) fused_steps_kernel(const uint32_t *walk_vtx, const uint64_t *walk_off, uint32_t n_walks, uint64_t n_steps,
                                                                    const uint4 *vinfo, uint32_t n_vtx, unsigned long long *tile_state, uint32_t *ticket, uint32_t *step_base,
                                                                    uint32_t *chunk_step, uint32_t *c_walk, uint64_t *walk_len, unsigned long long *ctr)
{
    __shared__ uint32_t sh_tile, sh_h[2];
    __shared__ uint64_t sh_ws, sh_excl, sh_warp[FS_THREADS / 32];
    constexpr int FS_TILE = FS_THREADS * FS_ITEMS;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) sh_tile = atomicAdd(ticket, 1u);              // tiles start in order: a tile only ever waits for tiles that run
    __syncthreads();
    const uint32_t tile = sh_tile;
    const uint64_t s0 = (uint64_t)tile * FS_TILE, s1 = min(s0 + FS_TILE, n_steps) - 1;
    if (wid < 2) {                                                      // walks of the two ends of the tile
        const uint32_t a = walk_of_step_warp(walk_off, n_walks, wid ? s1 : s0, lane);
        if (lane == 0) { sh_h[wid] = a; if (!wid) sh_ws = walk_off[a]; }
    }
    // the thread's steps and their vertex records (independent gathers), the record of the step before the first one
    const uint64_t sb = s0 + (uint64_t)FS_ITEMS * threadIdx.x;
    uint32_t vtx[FS_ITEMS];
    static_assert(FS_ITEMS % 4 == 0, "16-byte loads and stores per thread");
    if (sb + FS_ITEMS <= n_steps) {
        #pragma unroll
        for (int j = 0; j < FS_ITEMS; j += 4) { const uint4 q = *(const uint4 *)(walk_vtx + sb + j); vtx[j] = q.x; vtx[j + 1] = q.y; vtx[j + 2] = q.z; vtx[j + 3] = q.w; }
    } else { for (int j = 0; j < FS_ITEMS; ++j) vtx[j] = sb + j < n_steps ? walk_vtx[sb + j] : 0u; }
    uint4 me[FS_ITEMS];
    #pragma unroll
    for (int j = 0; j < FS_ITEMS; ++j) me[j] = vinfo_of(vinfo, n_vtx, vtx[j], ctr);
    uint32_t pb = __shfl_up_sync(0xFFFFFFFFu, me[FS_ITEMS - 1].y, 1), pt = __shfl_up_sync(0xFFFFFFFFu, me[FS_ITEMS - 1].z, 1);
    uint32_t pr = __shfl_up_sync(0xFFFFFFFFu, me[FS_ITEMS - 1].w, 1);
    if (lane == 0 && sb && sb < n_steps) { const uint4 p = vinfo_of(vinfo, n_vtx, walk_vtx[sb - 1], ctr); pb = p.y; pt = p.z; pr = p.w; }
    __syncthreads();
    const uint32_t h_lo = sh_h[0], h_hi = sh_h[1];
    // per step: first of its walk?  starts a chunk?  -> the thread's total
    uint32_t startm = 0, flagm = 0, zeros = 0; bool nonmono = false;
    uint64_t tot = 0;
    #pragma unroll
    for (int j = 0; j < FS_ITEMS; ++j) {
        const uint64_t s = sb + j;
        const bool valid = s < n_steps;
        uint64_t ws = sh_ws;
        if (h_lo != h_hi && valid) {
            uint32_t lo = h_lo, hi = h_hi + 1;
            while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (walk_off[m] <= s) lo = m; else hi = m; }
            ws = walk_off[lo];
        }
        const uint32_t qb = j ? me[j - 1 < 0 ? 0 : j - 1].y : pb, qt = j ? me[j - 1 < 0 ? 0 : j - 1].z : pt, qr = j ? me[j - 1 < 0 ? 0 : j - 1].w : pr;
        const bool start = valid && s == ws;
        const bool flag = valid && (start || me[j].y != qb || me[j].w != qr);
        if (valid && !start && (int32_t)qt >= (int32_t)me[j].z) nonmono = true;
        zeros += valid && me[j].x == 0;
        startm |= (start ? 1u : 0u) << j; flagm |= (flag ? 1u : 0u) << j;
        if (valid) tot = fs_comb(tot, (uint64_t)me[j].x | ((uint64_t)flag << 35) | (start ? FS_RESET : 0ull));
    }
    if (nonmono) ctr[CTR_NONMONO] = 1;
    const uint32_t bz = __ballot_sync(0xFFFFFFFFu, zeros != 0);
    if (bz) { for (int d = 16; d; d >>= 1) zeros += __shfl_xor_sync(0xFFFFFFFFu, zeros, d); if (lane == 0) atomicAdd(&ctr[CTR_ZERO_STEPS], (unsigned long long)zeros); }
    uint64_t inc = tot;                                                 // inclusive scan of the thread totals inside the warp
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint64_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc = fs_comb(o, inc); }
    if (lane == 31) sh_warp[wid] = inc;
    uint64_t tex = __shfl_up_sync(0xFFFFFFFFu, inc, 1); if (lane == 0) tex = 0;   // the threads of this warp before this one
    __syncthreads();
    if (wid == 0) {
        // exclusive prefixes of the warp totals, then the look-back: 32 predecessors per round, nearest in lane 0
        constexpr int NW = FS_THREADS / 32;
        uint64_t v = lane < NW ? sh_warp[lane] : 0ull;
        #pragma unroll
        for (int d = 1; d < NW; d <<= 1) { const uint64_t o = __shfl_up_sync(0xFFFFFFFFu, v, d); if (lane >= d) v = fs_comb(o, v); }
        const uint64_t total = __shfl_sync(0xFFFFFFFFu, v, NW - 1);
        uint64_t ex = __shfl_up_sync(0xFFFFFFFFu, v, 1); if (lane == 0) ex = 0;
        if (lane < NW) sh_warp[lane] = ex;
        uint64_t excl = 0;
        if (tile == 0) { if (lane == 0) ((volatile unsigned long long *)tile_state)[0] = (2ull << 62) | total; }
        else {
            if (lane == 0) ((volatile unsigned long long *)tile_state)[tile] = (1ull << 62) | total;
            uint64_t run = 0;                                              // aggregate of the tiles already folded in (nearer than the window)
            // 32 predecessors per round, nearest in lane 0.  (Looking at 8 predecessors per lane per round was tried — one latency for 256
            // tiles — and measured SLOWER, 8.7 ms instead of 6.4 ms on c4: a round then waits for the aggregate of every one of 256
            // concurrently running tiles instead of 32.)
            for (int64_t top = (int64_t)tile - 1;; top -= 32) {
                const int64_t j = top - lane;
                unsigned long long w = 2ull << 62;                          // before tile 0: an empty prefix
                if (j >= 0) { do { w = ((volatile unsigned long long *)tile_state)[j]; } while (!(w >> 62)); }
                const uint32_t pref = __ballot_sync(0xFFFFFFFFu, (w >> 62) == 2);
                const int stop = pref ? __ffs(pref) - 1 : 31;                // lanes 0..stop take part
                uint64_t v2 = lane <= stop ? (w & FS_VALUE) : 0ull;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint64_t o = __shfl_down_sync(0xFFFFFFFFu, v2, d); if (lane + d < 32) v2 = fs_comb(o, v2); }
                run = fs_comb(__shfl_sync(0xFFFFFFFFu, v2, 0), run);
                if (pref) break;
            }
            excl = run;
            if (lane == 0) ((volatile unsigned long long *)tile_state)[tile] = (2ull << 62) | fs_comb(excl, total);
        }
        if (lane == 0) sh_excl = excl;
    }
    __syncthreads();
    uint64_t pre = fs_comb(sh_excl, fs_comb(sh_warp[wid], tex));         // everything before the thread's first step
    uint32_t out[FS_ITEMS];
    #pragma unroll
    for (int j = 0; j < FS_ITEMS; ++j) {
        const uint64_t s = sb + j;
        const bool start = (startm >> j) & 1u, flag = (flagm >> j) & 1u;
        const uint64_t before = start ? 0ull : (pre & FS_BASES);
        out[j] = (uint32_t)before;
        if (s < n_steps) {
            uint32_t h = h_lo;
            if (h_lo != h_hi) {                                            // rare: the tile straddles a walk boundary
                uint32_t lo = h_lo, hi = h_hi + 1;
                while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (walk_off[m] <= s) lo = m; else hi = m; }
                h = lo;
            }
            const uint32_t c = (uint32_t)((pre & FS_CHUNKS) >> 35);
            if (flag) { chunk_step[c] = (uint32_t)s; c_walk[c] = h; }
            if (s + 1 == n_steps || s + 1 == walk_off[h + 1]) walk_len[h] = before + me[j].x;
            if (s + 1 == n_steps) chunk_step[c + (flag ? 1u : 0u)] = (uint32_t)n_steps;
            pre = fs_comb(pre, (uint64_t)me[j].x | ((uint64_t)flag << 35) | (start ? FS_RESET : 0ull));
        }
    }
    if (sb + FS_ITEMS <= n_steps) {
        #pragma unroll
        for (int j = 0; j < FS_ITEMS; j += 4) *(uint4 *)(step_base + sb + j) = make_uint4(out[j], out[j + 1], out[j + 2], out[j + 3]);
    } else { for (int j = 0; j < FS_ITEMS; ++j) if (sb + j < n_steps) step_base[sb + j] = out[j]; }
    // the chunk count is also kept by plain counting: it guards the 26-bit field of the scanned value
    uint32_t nf = __popc(flagm);
    #pragma unroll
    for (int d = 16; d; d >>= 1) nf += __shfl_xor_sync(0xFFFFFFFFu, nf, d);
    if (lane == 0 && nf) atomicAdd(&ctr[CTR_CHUNK_FLAGS], (unsigned long long)nf);
}

// scanned[s] = (chunks before s) << STEP_BASE_BITS | (bases before s, over all walks)
__global__ void __launch_bounds__(256) step_finalize_kernel(ChunkTable C, const PackedStep *packed, const uint64_t *scanned, const uint64_t *walk_off,
                                                            uint32_t n_walks, uint64_t n_steps, uint32_t *step_base)
{
    const uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (s == 0) C.chunk_step[C.n_chunks] = (uint32_t)n_steps;
    uint64_t ws;
    const uint64_t blk0 = blockIdx.x * (uint64_t)blockDim.x;
    const uint32_t h = walk_of_step_block(walk_off, n_walks, s < n_steps ? s : n_steps - 1, blk0, min(blk0 + blockDim.x, n_steps) - 1, ws);
    if (s >= n_steps) return;
    const uint64_t mask = (1ull << STEP_BASE_BITS) - 1;
    const uint64_t sc = scanned[s];
    step_base[s] = (uint32_t)((sc & mask) - (scanned[ws] & mask));
    if (packed[s].v >> 31) { const uint64_t c = sc >> STEP_BASE_BITS; if (c < C.n_chunks) { C.chunk_step[c] = (uint32_t)s; C.c_walk[c] = h; } }
}

__global__ void walk_len_kernel(const PackedStep *packed, const uint64_t *scanned, const uint64_t *walk_off, uint32_t n_walks, uint64_t n_steps, uint64_t *walk_len)
{
    uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_walks) return;
    const uint64_t mask = (1ull << STEP_BASE_BITS) - 1;
    const uint64_t end = n_steps ? (scanned[n_steps - 1] & mask) + (packed[n_steps - 1].v & 0x7FFFFFFFu) : 0;
    const uint64_t a = walk_off[h], b = walk_off[h + 1];
    walk_len[h] = (b < n_steps ? scanned[b] & mask : end) - (a < n_steps ? scanned[a] & mask : end);
}

// ---- zero-length steps (segments without bases contribute nothing, /root/reference/src/ILP_index.cpp:364-381): compaction
__global__ void nonzero_flags_kernel(const PackedStep *packed, uint64_t n, uint32_t *flags)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (packed[i].v & 0x7FFFFFFFu) != 0;
}
__global__ void compact_steps_kernel(const uint32_t *walk_vtx, const uint32_t *flags, const uint64_t *pos, uint64_t n, uint32_t *out_vtx)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n && flags[i]) out_vtx[pos[i]] = walk_vtx[i];
}
__global__ void remap_walk_off_kernel(const uint64_t *walk_off, uint32_t n_walks, const uint64_t *pos, uint64_t n_steps, uint64_t n_kept, uint64_t *out)
{
    uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h > n_walks) return;
    uint64_t o = walk_off[h];
    out[h] = o < n_steps ? pos[o] : n_kept;
}

// ---- geometry + fingerprint, 8 lanes per chunk (four chunks per warp: the chain of dependent loads of one chunk — walk, chunk
// bounds, step bases, context search — is latency, so the more chunks in flight the better)
// Owned windows: end positions e (start of the window's last k-mer) in [lo, hi), lo = first base of the chunk, hi = first base of
// the next chunk clipped to the walk's last k-mer.  Context steps [L, R]: from the step under base lo - w (halo window) to the
// step under base (next chunk start) + k - 2.
// A chunk whose first vertex lies outside the owned coordinate range (a walk region is set: another GPU sketches it) owns nothing.
constexpr int CK_LANES = 8;
__global__ void __launch_bounds__(256) chunk_key_kernel(ChunkTable C, const uint32_t *walk_vtx, const uint64_t *walk_off, const uint32_t *step_base,
                                                        const uint64_t *walk_len, const uint4 *vinfo, int k, int w, unsigned long long *ctr)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, sub = lane & (CK_LANES - 1), grp = lane / CK_LANES;
    const uint64_t c64 = (uint64_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 / CK_LANES) + grp;
    const bool live = c64 < C.n_chunks;                                  // lanes of a group beyond the last chunk only take part in the warp votes
    const uint32_t c = live ? (uint32_t)c64 : 0u;
    uint32_t h = 0; uint64_t ws = 0, we = 0, s0 = 0, s1 = 0; long long len = 0, lo = 0, b1 = 0, hi = 0, kpos = 0;
    if (live) {
        h = C.c_walk[c];
        ws = walk_off[h]; we = walk_off[h + 1];
        len = (long long)walk_len[h];
        s0 = C.chunk_step[c]; s1 = C.chunk_step[c + 1];
        lo = step_base[s0];
        b1 = s1 < we ? (long long)step_base[s1] : len;
        hi = min(b1, len - k + 1);
        const bool owned = vinfo[walk_vtx[s0]].w == 1u;
        kpos = (owned && len >= (long long)w + k - 1 && hi > lo) ? hi - lo : 0;   // k-mer positions of the walk that start in this chunk
        if (len < (long long)w + k - 1 || hi <= max(lo, (long long)w - 1) || !owned) hi = lo;     // no valid window ends here
    }
    const uint32_t gsh = CK_LANES * grp, gmask = (1u << CK_LANES) - 1u;
    uint32_t L = 0, R = 0;
    {   // L: going down from s0, the first step l with l == ws or step_base[l] <= lo - w (the group probes 8 candidate steps per round)
        const long long need = lo - w;
        uint64_t top = s0; bool done = !live;
        for (;;) {
            bool stop = false;
            if (!done) {
                const bool in = top >= ws + (uint64_t)sub;
                const uint64_t l = top - sub;
                stop = !in || l == ws || (long long)step_base[l] <= need;
            }
            const uint32_t b = (__ballot_sync(FULL, stop) >> gsh) & gmask;
            if (!done) { if (b) { L = (uint32_t)(top - (__ffs(b) - 1)); done = true; } else top -= CK_LANES; }
            if (__all_sync(FULL, done)) break;
        }
    }
    {   // R: going up from s1 - 1, the first step r with r + 1 == we or step_base[r + 1] > last
        const long long last = min(b1 + k - 2, len - 1);
        uint64_t bot = s1 - 1; bool done = !live;
        for (;;) {
            bool stop = false;
            if (!done) { const uint64_t r = bot + sub; stop = r + 1 >= we || (long long)step_base[r + 1] > last; }
            const uint32_t b = (__ballot_sync(FULL, stop) >> gsh) & gmask;
            if (!done) { if (b) { R = (uint32_t)(bot + (__ffs(b) - 1)); done = true; } else bot += CK_LANES; }
            if (__all_sync(FULL, done)) break;
        }
    }
    uint64_t h1 = 0, h2 = 0;
    if (live && hi > lo) {
        for (uint32_t i = L + sub; i <= R; i += CK_LANES) {
            const uint64_t x = walk_vtx[i], idx = i - L;
            const uint64_t m = cmix(((x + 1) << 32 | (idx + 1)) * 0x9E3779B97F4A7C15ull);
            h1 += m;
            h2 += (uint64_t)((uint32_t)(m >> 32) * 0x85EBCA6Bu + (uint32_t)idx) * (uint64_t)((uint32_t)m | 1u);   // a second, differently weighted sum
        }
    }
    #pragma unroll
    for (int d = CK_LANES / 2; d; d >>= 1) { h1 += __shfl_xor_sync(FULL, h1, d); h2 += __shfl_xor_sync(FULL, h2, d); }
    if (live && hi > lo) {
        const uint64_t meta = ((uint64_t)(s0 - L) << 40) ^ ((uint64_t)(s1 - L) << 20) ^ (uint64_t)(R - L);
        const uint64_t span = (uint64_t)(hi - lo);
        h1 = cmix(h1 ^ cmix(meta + 0x1234567ull) ^ (span << 32)); h2 = cmix(h2 + cmix(meta ^ 0xABCDEF01ull) + span);
    } else { h1 = h2 = 0; }
    // per warp: one atomic for the active chunks and one for the k-mer positions
    const bool lead = live && sub == 0;
    const uint32_t act = __ballot_sync(FULL, lead && hi > lo);
    unsigned long long kp = lead ? (unsigned long long)kpos : 0ull;
    #pragma unroll
    for (int d = 16; d; d >>= 1) kp += __shfl_xor_sync(FULL, kp, d);
    if (lane == 0) { if (act) atomicAdd(&ctr[CTR_ACTIVE_CHUNKS], (unsigned long long)__popc(act)); if (kp) atomicAdd(&ctr[CTR_PATH_POS], kp); }
    if (lead) { C.c_L[c] = L; C.c_R[c] = R; C.c_lo[c] = (uint32_t)lo; C.c_hi[c] = (uint32_t)hi; C.c_h1[c] = h1; C.c_h2[c] = h2; }
}

// ---- grouping by fingerprint: the slot ends up holding the smallest chunk id of its group
__global__ void chunk_group_kernel(ChunkTable C, uint32_t *table, uint32_t mask, int dedupe)
{
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C.n_chunks) return;
    if (C.c_hi[c] <= C.c_lo[c]) { C.c_slot[c] = C_NONE; return; }
    if (!dedupe) { C.c_slot[c] = C_NONE; return; }
    const uint64_t h1 = C.c_h1[c], h2 = C.c_h2[c];
    uint32_t slot = (uint32_t)h1 & mask;
    for (;;) {
        uint32_t cur = table[slot];
        if (cur == C_NONE) {
            uint32_t old = atomicCAS(&table[slot], C_NONE, c);
            if (old == C_NONE) break;
            cur = old;
        }
        if (C.c_h1[cur] == h1 && C.c_h2[cur] == h2) { atomicMin(&table[slot], c); break; }
        slot = (slot + 1) & mask;
    }
    C.c_slot[c] = slot;
}

// ---- representative of every chunk, member counts, exact verification (8 lanes per chunk)
__global__ void __launch_bounds__(256) chunk_rep_kernel(ChunkTable C, const uint32_t *table, const uint32_t *walk_vtx, int dedupe, int w, unsigned long long *ctr)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, sub = lane & (CK_LANES - 1), grp = lane / CK_LANES;
    const uint64_t c64 = (uint64_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 / CK_LANES) + grp;
    const bool live = c64 < C.n_chunks;
    const uint32_t c = live ? (uint32_t)c64 : 0u;
    const bool active = live && C.c_hi[c] > C.c_lo[c];
    uint32_t rep = C_NONE;
    if (active) rep = dedupe ? table[C.c_slot[c]] : c;
    bool same = true;
    if (active && rep != c) {
        // a member must match its representative exactly: same shape, same vertex sequence (the fingerprint only proposes)
        const uint32_t L = C.c_L[c], R = C.c_R[c], Lr = C.c_L[rep], Rr = C.c_R[rep];
        same = (R - L) == (Rr - Lr) && (C.chunk_step[c] - L) == (C.chunk_step[rep] - Lr) && (C.chunk_step[c + 1] - L) == (C.chunk_step[rep + 1] - Lr)
               && (C.c_hi[c] - C.c_lo[c]) == (C.c_hi[rep] - C.c_lo[rep]);
        if (same) for (uint32_t i = sub; i <= R - L; i += CK_LANES) same &= walk_vtx[L + i] == walk_vtx[Lr + i];
    }
    const uint32_t bad = (__ballot_sync(FULL, !same) >> (CK_LANES * grp)) & ((1u << CK_LANES) - 1u);
    if (live && sub == 0) {
        if (bad) ctr[CTR_DEDUPE_MISMATCH] = 1;
        C.c_rep[c] = rep;
        if (active) atomicAdd(&C.c_ninst[rep], 1u);
        const uint32_t T = (uint32_t)tile_cap(w, WALK_TILE_THREADS);
        const uint32_t nt = (active && rep == c) ? (C.c_hi[c] - C.c_lo[c] + T - 1) / T : 0u;
        C.c_ntile[c] = nt;
        if (nt) atomicAdd(&ctr[CTR_UNIQUE_WINDOWS], (unsigned long long)(C.c_hi[c] - C.c_lo[c]));
    }
}

// ---- tile records of the representatives (c_tile_base = exclusive scan of c_ntile)
__global__ void tile_fill_kernel(ChunkTable C, const uint64_t *walk_off, const uint64_t *walk_len, const uint32_t *step_base, int w, TileRec *tiles)
{
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C.n_chunks) return;
    const uint32_t nt = C.c_ntile[c];
    if (!nt) return;
    const uint32_t h = C.c_walk[c], lo = C.c_lo[c], hi = C.c_hi[c], R = C.c_R[c];
    const uint64_t ws = walk_off[h], we = walk_off[h + 1];
    const uint32_t wl = (uint32_t)walk_len[h];
    const uint32_t tb = C.c_tile_base[c];
    const uint32_t T = (hi - lo + nt - 1) / nt;                          // the chunk's windows, split evenly over its tiles (<= tile_cap each)
    for (uint32_t t = 0; t < nt; ++t) {
        TileRec r;
        r.walk = h; r.e0 = lo + t * T; r.e1 = min(r.e0 + T, hi); r.chunk = c; r.cbase = lo; r.walk_len = wl; r._r1 = 0;
        const long long first = max((long long)r.e0 - w - tile_pad(w), 0ll);   // first base the tile stages
        uint32_t a = (uint32_t)ws, b = R + 1;                           // last step with step_base <= first (the front padding may reach before L)
        while (b - a > 1) { uint32_t m = (a + b) >> 1; if ((long long)step_base[m] <= first) a = m; else b = m; }
        r.first_step = (uint32_t)(a - ws); r.step0 = a; r.step_end = we;
        tiles[tb + t] = r;
    }
}

// ---- per-walk minimizer counts: every member adds what its representative emitted
__global__ void chunk_emitted_kernel(ChunkTable C, const uint32_t *c_emitted, const uint32_t *c_hits, uint32_t walk_id_base,
                                     unsigned long long *minimizers_per_walk, unsigned long long *ctr)
{
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t h = C_NONE; unsigned long long v = 0, nh = 0;
    if (c < C.n_chunks) { uint32_t rep = C.c_rep[c]; if (rep != C_NONE) { h = C.c_walk[c]; v = c_emitted[rep]; nh = c_hits[rep]; } }
    #pragma unroll
    for (int d = 16; d; d >>= 1) nh += __shfl_xor_sync(0xFFFFFFFFu, nh, d);
    if ((threadIdx.x & 31) == 0 && nh) atomicAdd(&ctr[CTR_PATH_HITS], nh);
    // chunks are ordered by walk: lanes of one warp mostly share the walk -> one atomic per distinct walk per warp
    uint32_t peers = __match_any_sync(0xFFFFFFFFu, h);
    unsigned long long sum = 0;
    for (uint32_t m = peers; m; m &= m - 1) sum += __shfl_sync(peers, v, __ffs(m) - 1);
    if (h != C_NONE && (threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1) && sum) atomicAdd(&minimizers_per_walk[walk_id_base + h], sum);
}

// ---- survivors per representative chunk: hits of its tiles whose rank is not dropped (one warp per tile)
__global__ void __launch_bounds__(256) tile_survivors_kernel(const TileRec *tiles, uint32_t n_tiles, const uint32_t *seg_off, const uint32_t *seg_cnt,
                                                             const uint32_t *hit_rank, const uint8_t *hit_nv, const uint8_t *rank_drop,
                                                             uint32_t *c_surv, uint32_t *c_surv_vtx)
{
    const uint32_t tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (tile >= n_tiles) return;
    uint32_t n = 0, nv = 0;
    for (int sg = 0; sg < SEG_PER_TILE; ++sg) {
        const uint32_t cnt = seg_cnt[(size_t)tile * SEG_PER_TILE + sg], off = seg_off[(size_t)tile * SEG_PER_TILE + sg];
        for (uint32_t i = lane; i < cnt; i += 32) if (!rank_drop[hit_rank[off + i]]) { ++n; nv += hit_nv[off + i]; }
    }
    #pragma unroll
    for (int d = 16; d; d >>= 1) { n += __shfl_xor_sync(0xFFFFFFFFu, n, d); nv += __shfl_xor_sync(0xFFFFFFFFu, nv, d); }
    if (lane == 0 && n) { atomicAdd(&c_surv[tiles[tile].chunk], n); atomicAdd(&c_surv_vtx[tiles[tile].chunk], nv); }
}
// records (and their vertices, summed into ctr[CTR_SURV_VTX]) every member chunk will instantiate
__global__ void member_counts_kernel(ChunkTable C, const uint32_t *c_surv, const uint32_t *c_surv_vtx, uint32_t *member_cnt, unsigned long long *ctr)
{
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    if (c < C.n_chunks) {
        uint32_t rep = C.c_rep[c];
        member_cnt[c] = rep == C_NONE ? 0u : c_surv[rep];
        v = rep == C_NONE ? 0u : c_surv_vtx[rep];
    }
    #pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&ctr[CTR_SURV_VTX], v);
}

// ---- instantiate the surviving hits of every member chunk (one warp per member), in (walk, position) order:
// members are numbered along their walks, tiles along the chunk, hits along the tile
__global__ void __launch_bounds__(256) expand_kernel(ChunkTable C, ExpandArgs X)
{
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= C.n_chunks) return;
    const uint32_t rep = C.c_rep[c];
    if (rep == C_NONE) return;
    uint64_t out = X.member_off[c];
    if (X.member_off[c + 1] == out) return;
    const uint32_t walk = X.walk_id_base + C.c_walk[c];
    const int32_t shift = (int32_t)C.c_lo[c] - X.w;                          // rel = p + w - lo(rep)  ->  p' = rel - w + lo(member)
    const uint32_t t0 = C.c_tile_base[rep], nt = C.c_ntile[rep];
    for (uint32_t t = t0; t < t0 + nt; ++t) {
        for (int sg = 0; sg < SEG_PER_TILE; ++sg) {
            const uint32_t cnt = X.hseg_cnt[(size_t)t * SEG_PER_TILE + sg], off = X.hseg_off[(size_t)t * SEG_PER_TILE + sg];
            for (uint32_t i0 = 0; i0 < cnt; i0 += 32) {
                const uint32_t i = off + i0 + lane;
                const bool have = i0 + lane < cnt;
                uint32_t r = have ? X.hit_rank[i] : 0u;
                const bool keep = have && !X.rank_drop[r];
                const uint32_t b = __ballot_sync(0xFFFFFFFFu, keep);
                if (keep) {
                    const uint64_t j = out + __popc(b & lanemask_lt());
                    X.x_rank[j] = r; X.x_walk[j] = walk; X.x_pos[j] = (uint32_t)((int32_t)X.hit_pos[i] + shift);
                    X.x_voff[j] = X.hit_voff[i]; X.x_nv[j] = X.hit_nv[i];
                    if (X.x_hash) X.x_hash[j] = X.hit_hash[i];
                }
                out += __popc(b);
            }
        }
    }
}

// ------------------------------------------------------------------ launchers
cudaError_t chunk_topo_coord(const int32_t *top_order_map, const uint64_t *seg_off, uint32_t n_vtx, int shift, uint64_t own_lo, uint64_t own_hi,
                             uint32_t *tlen, uint64_t *prefix, uint4 *vinfo, void *scan_scratch, unsigned long long *ctr, cudaStream_t st, uint64_t *launches)
{
    if (!n_vtx) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(tlen, 0, (size_t)n_vtx * 4, st);
    if (e != cudaSuccess) return e;
    topo_len_kernel<<<(n_vtx + 255) / 256, 256, 0, st>>>(top_order_map, seg_off, n_vtx, tlen, ctr);
    PHI_LAUNCH_CHECK();
    e = scan_u32_to_u64(tlen, prefix, n_vtx, scan_scratch, st, launches);
    if (e != cudaSuccess) return e;
    topo_coord_kernel<<<(n_vtx + 255) / 256, 256, 0, st>>>(top_order_map, seg_off, prefix, n_vtx, shift, own_lo, own_hi, ctr, vinfo);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t walk_step_pass(const uint32_t *walk_vtx, const uint64_t *walk_off, uint32_t n_walks, uint64_t n_steps, const uint4 *vinfo, uint32_t n_vtx,
                           PackedStep *packed, unsigned long long *ctr, cudaStream_t st, uint64_t *launches)
{
    if (!n_steps) return cudaSuccess;
    step_pass_kernel<<<(unsigned)((n_steps + 255) / 256), 256, 0, st>>>(walk_vtx, walk_off, n_walks, n_steps, vinfo, n_vtx, packed, ctr);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

static int fs_threads()
{
    static int v = 0;
    if (!v) { const char *e = getenv("PHI_GPU_FS_THREADS"); v = e ? atoi(e) : 256; if (v != 128 && v != 256) v = 256; }
    return v;
}
static int fs_items()
{
    static int v = 0;
    if (!v) { const char *e = getenv("PHI_GPU_FS_ITEMS"); v = e ? atoi(e) : 8; if (v != 4 && v != 8 && v != 16) v = 8; }
    return v;
}

cudaError_t walk_steps_fused(const uint32_t *walk_vtx, const uint64_t *walk_off, uint32_t n_walks, uint64_t n_steps, const uint4 *vinfo, uint32_t n_vtx,
                             unsigned long long *tile_state, uint32_t *ticket, uint32_t *step_base, uint32_t *chunk_step, uint32_t *c_walk,
                             uint64_t *walk_len, unsigned long long *ctr, cudaStream_t st, uint64_t *launches, uint64_t tile_first, uint64_t tile_last)
{
    if (!n_steps) return cudaSuccess;
    const uint64_t nt = walk_steps_fused_tiles(n_steps);
    if (tile_last > nt) tile_last = nt;
    if (tile_first == 0) {
        cudaError_t e = cudaMemsetAsync(tile_state, 0, nt * 8, st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(ticket, 0, 4, st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(walk_len, 0, (size_t)n_walks * 8, st);       // walks without steps
        if (e != cudaSuccess) return e;
    }
    if (tile_last <= tile_first) return cudaSuccess;
    const unsigned grid = (unsigned)(tile_last - tile_first);
#define PHI_FS_LAUNCH(I, T) fused_steps_kernel<I, T><<<grid, T, 0, st>>>(walk_vtx, walk_off, n_walks, n_steps, vinfo, n_vtx, tile_state, ticket, step_base, chunk_step, c_walk, walk_len, ctr)
    if (fs_threads() == 128) { switch (fs_items()) { case 4: PHI_FS_LAUNCH(4, 128); break; case 16: PHI_FS_LAUNCH(16, 128); break; default: PHI_FS_LAUNCH(8, 128); break; } }
    else { switch (fs_items()) { case 4: PHI_FS_LAUNCH(4, 256); break; case 16: PHI_FS_LAUNCH(16, 256); break; default: PHI_FS_LAUNCH(8, 256); break; } }
#undef PHI_FS_LAUNCH
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}
uint64_t walk_steps_fused_tile_steps() { return (uint64_t)fs_threads() * fs_items(); }
uint64_t walk_steps_fused_tiles(uint64_t n_steps) { const uint64_t t = walk_steps_fused_tile_steps(); return (n_steps + t - 1) / t; }

cudaError_t walk_step_finalize(const ChunkTable &C, const PackedStep *packed, const uint64_t *scanned, const uint64_t *walk_off, uint32_t n_walks,
                               uint64_t n_steps, uint32_t *step_base, uint64_t *walk_len, cudaStream_t st, uint64_t *launches)
{
    if (n_steps) {
        step_finalize_kernel<<<(unsigned)((n_steps + 255) / 256), 256, 0, st>>>(C, packed, scanned, walk_off, n_walks, n_steps, step_base);
        PHI_LAUNCH_CHECK();
    }
    if (n_walks) {
        walk_len_kernel<<<(n_walks + 127) / 128, 128, 0, st>>>(packed, scanned, walk_off, n_walks, n_steps, walk_len);
        PHI_LAUNCH_CHECK();
    }
    return cudaSuccess;
}

cudaError_t walk_compact_steps(const uint32_t *walk_vtx, const uint64_t *walk_off, uint32_t n_walks, uint64_t n_steps, const PackedStep *packed,
                               uint32_t *flags, uint64_t *pos, void *scan_scratch, uint64_t n_kept, uint32_t *out_vtx, uint64_t *out_off,
                               cudaStream_t st, uint64_t *launches)
{
    if (!n_steps) return cudaSuccess;
    nonzero_flags_kernel<<<(unsigned)((n_steps + 255) / 256), 256, 0, st>>>(packed, n_steps, flags);
    PHI_LAUNCH_CHECK();
    cudaError_t e = scan_u32_to_u64(flags, pos, n_steps, scan_scratch, st, launches);
    if (e != cudaSuccess) return e;
    compact_steps_kernel<<<(unsigned)((n_steps + 255) / 256), 256, 0, st>>>(walk_vtx, flags, pos, n_steps, out_vtx);
    PHI_LAUNCH_CHECK();
    remap_walk_off_kernel<<<(n_walks + 1 + 127) / 128, 128, 0, st>>>(walk_off, n_walks, pos, n_steps, n_kept, out_off);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t chunk_keys(const ChunkTable &C, const uint32_t *walk_vtx, const uint64_t *walk_off, const uint32_t *step_base, const uint64_t *walk_len,
                       const uint4 *vinfo, int k, int w, unsigned long long *ctr, cudaStream_t st, uint64_t *launches)
{
    if (!C.n_chunks) return cudaSuccess;
    chunk_key_kernel<<<(unsigned)(((uint64_t)C.n_chunks * CK_LANES + 255) / 256), 256, 0, st>>>(C, walk_vtx, walk_off, step_base, walk_len, vinfo, k, w, ctr);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t chunk_group(const ChunkTable &C, uint32_t *table, uint32_t table_cap, const uint32_t *walk_vtx, int dedupe, int w, unsigned long long *ctr,
                        cudaStream_t st, uint64_t *launches)
{
    if (!C.n_chunks) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(C.c_ninst, 0, (size_t)C.n_chunks * 4, st);
    if (e != cudaSuccess) return e;
    if (dedupe) {
        e = fill_u32(table, table_cap, C_NONE, st, launches);
        if (e != cudaSuccess) return e;
    }
    chunk_group_kernel<<<(C.n_chunks + 255) / 256, 256, 0, st>>>(C, table, table_cap - 1, dedupe);
    PHI_LAUNCH_CHECK();
    chunk_rep_kernel<<<(unsigned)(((uint64_t)C.n_chunks * CK_LANES + 255) / 256), 256, 0, st>>>(C, table, walk_vtx, dedupe, w, ctr);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t chunk_tiles(const ChunkTable &C, const uint64_t *walk_off, const uint64_t *walk_len, const uint32_t *step_base, int w, TileRec *tiles, cudaStream_t st, uint64_t *launches)
{
    if (!C.n_chunks) return cudaSuccess;
    tile_fill_kernel<<<(C.n_chunks + 127) / 128, 128, 0, st>>>(C, walk_off, walk_len, step_base, w, tiles);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t chunk_emitted(const ChunkTable &C, const uint32_t *c_emitted, const uint32_t *c_hits, uint32_t walk_id_base,
                          unsigned long long *minimizers_per_walk, unsigned long long *ctr, cudaStream_t st, uint64_t *launches)
{
    if (!C.n_chunks) return cudaSuccess;
    chunk_emitted_kernel<<<(C.n_chunks + 255) / 256, 256, 0, st>>>(C, c_emitted, c_hits, walk_id_base, minimizers_per_walk, ctr);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t chunk_survivors(const ChunkTable &C, const TileRec *tiles, uint32_t n_tiles, const uint32_t *seg_off, const uint32_t *seg_cnt,
                            const uint32_t *hit_rank, const uint8_t *hit_nv, const uint8_t *rank_drop, uint32_t *c_surv, uint32_t *c_surv_vtx,
                            uint32_t *member_cnt, unsigned long long *ctr, cudaStream_t st, uint64_t *launches)
{
    if (!C.n_chunks) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(c_surv, 0, (size_t)C.n_chunks * 4, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(c_surv_vtx, 0, (size_t)C.n_chunks * 4, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(ctr + CTR_SURV_VTX, 0, 8, st);
    if (e != cudaSuccess) return e;
    if (n_tiles) {
        tile_survivors_kernel<<<(unsigned)(((uint64_t)n_tiles * 32 + 255) / 256), 256, 0, st>>>(tiles, n_tiles, seg_off, seg_cnt, hit_rank, hit_nv, rank_drop, c_surv, c_surv_vtx);
        PHI_LAUNCH_CHECK();
    }
    member_counts_kernel<<<(C.n_chunks + 255) / 256, 256, 0, st>>>(C, c_surv, c_surv_vtx, member_cnt, ctr);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t chunk_expand(const ChunkTable &C, const ExpandArgs &X, cudaStream_t st, uint64_t *launches)
{
    if (!C.n_chunks) return cudaSuccess;
    expand_kernel<<<(unsigned)(((uint64_t)C.n_chunks * 32 + 255) / 256), 256, 0, st>>>(C, X);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

}  // namespace phi
