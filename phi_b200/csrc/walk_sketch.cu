// Walk-sketch kernel (hand-written sm_100a CUDA), 128-thread tiles.
//
//   walk_sketch_kernel : ILP_index::index_kmers for the representative chunks of the walks + compute_anchors
//                        (/root/reference/src/ILP_index.cpp:359-445, :495-526, :643-655), fused:
//                        node-spanning windows are gathered into shared memory from the segment
//                        store, minimizers are probed against the ranked read spectrum as they are
//                        found, and only hits (rank, chunk, position, vertex list) are written.
#define PHI_TILE_THREADS 128
#include "kernels.h"
#include "sketch_common.cuh"
#include <cstdlib>

namespace phi {
static_assert(PHI_TILE_THREADS == WALK_TILE_THREADS, "tile size");

// ================================================================== walks
// rank of `key` in the sorted spectrum via the radix directory over the top dbits bits; -1 if absent
__device__ __forceinline__ int64_t spectrum_probe(const uint64_t *spec, const uint32_t *dir, int dbits, uint64_t key)
{
    uint32_t b = dbits ? (uint32_t)(key >> (64 - dbits)) : 0u;
    uint32_t lo = dir[b], hi = dir[b + 1];
    for (uint32_t i = lo; i < hi; ++i) {
        uint64_t s = spec[i];
        if (s == key) return (int64_t)i;
        if (s > key) break;
    }
    return -1;
}

// Slow anchor path: distinct vertices in first-seen order, then sorted by top_order_map
// (/root/reference/src/ILP_index.cpp:424-435).  Returns the count; writes the list if out != nullptr.
__device__ __noinline__ int anchor_slow(const Tile &t, int j0, int n_raw, const int32_t *top_order_map, int32_t *out)
{
    int32_t u[MAX_K]; int nu = 0;
    for (int i = 0; i < n_raw; ++i) {
        int32_t v = (int32_t)t.stepv[j0 + i];
        bool seen = false;
        for (int q = 0; q < nu; ++q) seen |= (u[q] == v);
        if (!seen) u[nu++] = v;
    }
    for (int a = 1; a < nu; ++a) {
        int32_t x = u[a]; int32_t tx = top_order_map[x]; int b = a - 1;
        while (b >= 0 && top_order_map[u[b]] > tx) { u[b + 1] = u[b]; --b; }
        u[b + 1] = x;
    }
    if (out) for (int i = 0; i < nu; ++i) out[i] = u[i];
    return nu;
}

// step index of local base position p (chunk directory filled while gathering)
__device__ __forceinline__ int step_of(const Tile &t, int p)
{
    const int c = p >> 3;
    return t.cfirst[c] + __popc((uint32_t)t.cmask[c] & ((2u << (p & 7)) - 1u));
}

template <bool CLEAN, bool FAST>
__device__ __forceinline__ void walk_tile_body(Tile &t, const WalkSketchArgs &A, const TileRec &tr, uint32_t tile)
{
    const int tid = threadIdx.x;
    uint16_t *runs = t.pre;
    uint64_t *run_val = t.canon;                                     // FAST only
    int halo = -1, n_runs; uint64_t halo_val = 0;
    if (FAST) {
        n_runs = fast_runs<false>(t, runs, run_val, &halo, &halo_val);
    } else {
        phase_canon<CLEAN>(t);
        __syncthreads();
        phase_block_minima<CLEAN>(t);
        __syncthreads();
        n_runs = phase_runs<false, CLEAN>(t, runs, &halo);
    }
    if (n_runs == 0) return;

    if (FAST) { if (tid == tile_halo_lanes(t.w)) t.hash[0] = halo >= 0 ? hash_packed_kmer(halo_val, t.k) : 0xFFFFFFFFFFFFFFFFull; }
    else if (tid == 0) t.hash[0] = halo >= 0 ? hash_at<CLEAN>(t, halo) : 0xFFFFFFFFFFFFFFFFull;
    int emitted = 0;
    for (int b0 = 0; b0 < n_runs; b0 += NT) {
        const int jr = b0 + tid; const bool have = jr < n_runs;
        const int cnt = min(NT, n_runs - b0);
        uint32_t ent = have ? runs[jr] : 0;
        const int a = ent & 0x7FFF;
        uint64_t hv = 0;
        if (have) hv = FAST ? hash_packed_kmer(run_val[jr], t.k) : hash_at<CLEAN>(t, a);
        t.hash[tid + 1] = hv;
        __syncthreads();
        uint64_t prev = (ent & 0x8000) ? 0xFFFFFFFFFFFFFFFFull : t.hash[tid];
        uint64_t carry = t.hash[cnt];
        bool emit = have && hv != prev;
        emitted += __syncthreads_count(emit);
        if (tid == 0) t.hash[0] = carry;

        // ---- probe + anchor
        int64_t rank = -1;
        if (emit) rank = A.mode == WALK_MODE_ALL ? 0 : spectrum_probe(A.spec, A.dir, A.dbits, hv);
        const bool hit = rank >= 0;
        int j0 = 0, nv = 0, n_raw = 0; bool slow = false;
        if (hit) {
            j0 = step_of(t, a);
            n_raw = nv = step_of(t, a + A.k - 1) - j0 + 1;           // steps under bases [a, a+k)
            if (!A.walks_monotone) {                                 // walk order == topological order for a valid walk; verify otherwise
                int32_t prev_top = -0x7FFFFFFF - 1;
                for (int i = 0; i < nv; ++i) {
                    int32_t tp = A.top_order_map[t.stepv[j0 + i]];
                    if (i && tp <= prev_top) slow = true;
                    prev_top = tp;
                }
                if (slow) nv = anchor_slow(t, j0, n_raw, A.top_order_map, nullptr);
            }
        }
        Scan2 sc = block_scan2(t.scan, hit ? 1 : 0, nv);
        __shared__ unsigned long long s_base_hit, s_base_vtx;
        if (tid == 0 && sc.tot_a) {
            s_base_hit = atomicAdd(&A.ctr[CTR_HITS], (unsigned long long)sc.tot_a);
            s_base_vtx = atomicAdd(&A.ctr[CTR_HIT_VTX], (unsigned long long)sc.tot_b);
            if (s_base_hit + sc.tot_a <= A.hit_cap) {                // this batch's hits: one contiguous segment, in position order
                A.hseg_off[(size_t)tile * SEG_PER_TILE + b0 / NT] = (uint32_t)s_base_hit;
                A.hseg_cnt[(size_t)tile * SEG_PER_TILE + b0 / NT] = (uint32_t)sc.tot_a;
            }
            atomicAdd(&A.chunk_hits[tr.chunk], (uint32_t)sc.tot_a);
        }
        __syncthreads();
        if (hit) {
            unsigned long long hi_idx = s_base_hit + sc.ex_a, vo = s_base_vtx + sc.ex_b;
            if (hi_idx < A.hit_cap && vo + nv <= A.vtx_cap) {
                A.hit_rank[hi_idx] = (uint32_t)rank;
                A.hit_chunk[hi_idx] = tr.chunk;
                A.hit_pos[hi_idx] = (uint32_t)(t.g0 + a + A.w - (long long)tr.cbase);
                A.hit_voff[hi_idx] = vo;
                A.hit_nv[hi_idx] = (uint8_t)nv;
                if (A.hit_hash) A.hit_hash[hi_idx] = hv;
                if (!slow) for (int i = 0; i < nv; ++i) A.vtx_pool[vo + i] = (int32_t)t.stepv[j0 + i];
                else anchor_slow(t, j0, n_raw, A.top_order_map, A.vtx_pool + vo);
                if (A.probe) {                                       // (rank, vertices, first 6 vertices): what the group table compares
                    uint32_t v[6];
                    #pragma unroll
                    for (int q = 0; q < 6; ++q) v[q] = q < nv ? (slow ? (uint32_t)A.vtx_pool[vo + q] : t.stepv[j0 + q]) : 0u;
                    A.probe[2 * hi_idx] = make_uint4((uint32_t)rank, (uint32_t)nv, v[0], v[1]);
                    A.probe[2 * hi_idx + 1] = make_uint4(v[2], v[3], v[4], v[5]);
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0 && emitted) atomicAdd(&A.chunk_emitted[tr.chunk], (uint32_t)emitted);
}

__device__ __forceinline__ void walk_sketch_body(const WalkSketchArgs &A)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t tile = blockIdx.x;
    const TileRec tr = A.tiles[tile];
    const long long len = tr.walk_len;
    const TileLayout &L = A.layout;
    Tile t = carve(smem, L, A.k, A.w);
    // the tile owns the window end positions [e0, e1) of walk h (at most TILE_W); everything below is sized by what it really holds
    t.M = (int)(tr.e1 - tr.e0) + A.w + L.pad; t.M8 = (t.M + 7) & ~7; t.NB = t.M + A.k - 1;
    const int nchunks = (t.NB + 7) >> 3;
    t.g0 = (long long)tr.e0 - A.w - L.pad;
    t.seq_len = len;
    set_window_bounds(t);
    const int tid = threadIdx.x;

    // ---- steps overlapping the tile's bases [base_lo, base_hi)
    const long long base_lo = t.g0 < 0 ? 0 : t.g0;
    const long long base_hi = min(len, t.g0 + (long long)t.NB);
    const uint64_t wend = tr.step_end, s0 = tr.step0;
    int n_steps = 0;
    for (uint64_t c0 = s0;; c0 += NT) {
        uint64_t s = c0 + tid; int ok = 0;
        if (s < wend) {
            long long sb = A.step_base[s];
            if (sb < base_hi) {
                ok = 1;
                int j = (int)(s - s0);
                t.stepv[j] = A.walk_vtx[s];
                t.steps[j] = (uint16_t)(sb <= base_lo ? 0 : sb - base_lo);
            }
        }
        int c = __syncthreads_count(ok);
        n_steps += c;
        if (c < NT) break;
    }
    if (tid == 0) t.steps[n_steps] = (uint16_t)(base_hi - base_lo);
    const long long first_true = A.step_base[s0];                    // true start of step 0 (may precede base_lo)
    __syncthreads();

    // ---- gather bases through the step table: per 8-base chunk, one unaligned 8-byte load per overlapping step
    uint32_t dirty_any = 0;
    const int rel0 = (int)(base_lo - t.g0);                          // local index of base_lo (0, or w for tile 0)
    const int nb = (int)(base_hi - base_lo);                         // real bases staged
    for (int c = tid; c < nchunks; c += NT) {
        const int q0 = 8 * c - rel0;                                 // chunk start relative to base_lo (may be < 0)
        uint64_t v = 0; int first = 0; uint32_t smask = 0;
        if (q0 + 8 > 0 && q0 < nb) {
            int cur = q0 < 0 ? 0 : q0;
            const int qend = min(q0 + 8, nb);
            int a = 0, b = n_steps;                                  // last j with steps[j] <= cur
            while (b - a > 1) { int m = (a + b) >> 1; if (t.steps[m] <= cur) a = m; else b = m; }
            int j = first = a;
            while (cur < qend) {
                const int jend = t.steps[j + 1];
                const int hi = min(jend, qend);
                const long long jstart = j == 0 ? first_true - base_lo : (long long)t.steps[j];   // relative to base_lo, may be < 0 for j == 0
                const uint8_t *src = A.seg_bases + A.seg_off[t.stepv[j]] + ((long long)q0 - jstart);  // chunk byte 0 in this segment's coordinates
                uint64_t x = load8_unaligned(src);
                const int lo_b = cur - q0, hi_b = hi - q0;           // bytes [lo_b, hi_b) of the chunk come from step j
                uint64_t m = (hi_b == 8 ? ~0ull : ((1ull << (8 * hi_b)) - 1)) & (~0ull << (8 * lo_b));
                v |= x & m;
                if (j != first) smask |= 1u << lo_b;
                cur = hi;
                if (cur == jend) ++j;
            }
            v = upcase8(v);
        }
        t.cfirst[c] = (uint16_t)first; t.cmask[c] = (uint8_t)smask;
        dirty_any |= stage_chunk(t, c, v);
    }
    if (__syncthreads_or(dirty_any != 0 || A.k > MAX_PACKED_K)) walk_tile_body<false, false>(t, A, tr, tile);
    else if (tile_fast_w(A.w)) walk_tile_body<true, true>(t, A, tr, tile);
    else walk_tile_body<true, false>(t, A, tr, tile);
}

// The same body under three register budgets (resident CTAs per SM: 6 -> 80 registers, 7 -> 72, 8 -> 64 with a few spills); which one
// wins is a measurement (PHI_GPU_WALK_CTAS picks; default below).
__global__ void __launch_bounds__(NT, 6) walk_sketch_kernel(WalkSketchArgs A) { walk_sketch_body(A); }
__global__ void __launch_bounds__(NT, 7) walk_sketch_kernel_r72(WalkSketchArgs A) { walk_sketch_body(A); }
__global__ void __launch_bounds__(NT, 8) walk_sketch_kernel_r64(WalkSketchArgs A) { walk_sketch_body(A); }

// ================================================================== hash KAT hook
__global__ void hash_bytes_kernel(const uint8_t *keys, uint64_t n, int len, uint64_t *out)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *p = keys + i * (uint64_t)len;
    bool clean = true;
    for (int j = 0; j < len; ++j) clean &= is_acgt(p[j]);
    uint64_t h = murmur3_x64_128_xor_bytes([&](int j) -> uint32_t { return p[j]; }, len);
    if (len <= 32) {                                                  // the word form used for short keys must agree
        uint64_t W[4] = {0, 0, 0, 0};
        for (int j = 0; j < len; ++j) W[j >> 3] |= (uint64_t)p[j] << (8 * (j & 7));
        if (murmur3_x64_128_xor(W, len) != h) h = ~h;
    }
    if (clean && len <= 32) {                                         // cross-check the packed path used by the sketch kernels
        uint64_t km = 0;
        for (int j = 0; j < len; ++j) km = (km << 2) | code2(p[j]);
        uint64_t h2 = hash_packed_kmer(km, len);
        if (h2 != h) h = ~h;                                          // make any divergence visible to the test
    }
    out[i] = h;
}

// ------------------------------------------------------------------ launchers

cudaError_t launch_walk_sketch(const WalkSketchArgs &A, uint32_t n_tiles, cudaStream_t st)
{
    if (!n_tiles) return cudaSuccess;
    size_t smem = (size_t)A.layout.bytes;
    static int ctas = 0;
    if (!ctas) { const char *e = getenv("PHI_GPU_WALK_CTAS"); ctas = e ? atoi(e) : 8; if (ctas < 6 || ctas > 8) ctas = 8; }
    auto kern = ctas == 8 ? walk_sketch_kernel_r64 : ctas == 7 ? walk_sketch_kernel_r72 : walk_sketch_kernel;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<n_tiles, NT, smem, st>>>(A);
    return cudaGetLastError();
}

cudaError_t launch_hash_bytes(const uint8_t *keys, uint64_t n, int len, uint64_t *out, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    hash_bytes_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(keys, n, len, out);
    return cudaGetLastError();
}

TileLayout walk_tile_layout(int k, int w) { return make_layout(k, w, true); }

}  // namespace phi
