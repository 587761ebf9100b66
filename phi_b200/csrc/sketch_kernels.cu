// Read-sketch and walk-sketch kernels (hand-written sm_100a CUDA).
//
//   read_sketch_kernel : ILP_index::compute_hashes for every read + the Sp_R union
//                        (/root/reference/src/ILP_index.cpp:447-493, :615-629), fused:
//                        minimizer hashes go straight into an open-addressing HBM table.
//   walk_sketch_kernel : ILP_index::index_kmers for every walk + compute_anchors
//                        (/root/reference/src/ILP_index.cpp:359-445, :495-526, :643-655), fused:
//                        node-spanning windows are gathered into shared memory from the segment
//                        store, minimizers are probed against the ranked read spectrum as they are
//                        found, and only hits (rank, walk, position, vertex list) are written.
#include "kernels.h"
#include "sketch_tile.cuh"

namespace phi {

// ------------------------------------------------------------------ block scan of two ints
struct Scan2 { int ex_a, ex_b, tot_a, tot_b; };
__device__ __forceinline__ Scan2 block_scan2(uint32_t *scratch /* >= 32 words */, int a, int b)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int ia = a, ib = b;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int ta = __shfl_up_sync(0xFFFFFFFFu, ia, d), tb = __shfl_up_sync(0xFFFFFFFFu, ib, d);
        if (lane >= d) { ia += ta; ib += tb; }
    }
    if (lane == 31) { scratch[2 * wid] = ia; scratch[2 * wid + 1] = ib; }
    __syncthreads();
    Scan2 s; s.ex_a = ia - a; s.ex_b = ib - b; s.tot_a = 0; s.tot_b = 0;
    #pragma unroll
    for (int i = 0; i < NT / 32; ++i) {
        int ca = scratch[2 * i], cb = scratch[2 * i + 1];
        if (i < wid) { s.ex_a += ca; s.ex_b += cb; }
        s.tot_a += ca; s.tot_b += cb;
    }
    __syncthreads();
    return s;
}

// ================================================================== reads
// Insert into the ORDER-PRESERVING open-addressing spectrum table: the home slot is a monotone function of the key
// (top bits of the hash: umulhi(key, mult)), collisions probe upwards without wrap-around.  Keys therefore end up sorted
// at the granularity of probe clusters (runs of occupied slots), and sorting each short cluster in place
// (primitives.cu: table_sort_clusters) leaves the whole table in ascending order — no radix sort of the spectrum.
// u64 keys, EMPTY = ~0; the key ~0 itself is recorded in ctr[CTR_HAS_MAXKEY] instead of the table.
__device__ __forceinline__ void table_insert(uint64_t *table, uint64_t mult, uint64_t limit, uint64_t key, unsigned long long *ctr)
{
    if (key == TABLE_EMPTY) { ctr[CTR_HAS_MAXKEY] = 1; return; }
    for (uint64_t slot = __umul64hi(key, mult); slot < limit; ++slot) {
        uint64_t cur = table[slot];
        if (cur == key) return;
        if (cur == TABLE_EMPTY) {
            uint64_t old = atomicCAS((unsigned long long *)&table[slot], (unsigned long long)TABLE_EMPTY, (unsigned long long)key);
            if (old == TABLE_EMPTY || old == key) return;
        }
    }
    ctr[CTR_OVERFLOW] = 1;                                           // ran off the padding behind the last home slot: the host retries larger
}

// unaligned 8-byte load (two aligned loads + funnel); the buffers are padded so that p-7 .. p+15 is always readable
__device__ __forceinline__ uint64_t load8_unaligned(const uint8_t *p)
{
    const unsigned long long a = (unsigned long long)p;
    const uint64_t *w = (const uint64_t *)(a & ~7ull);
    const int s = (int)(a & 7) * 8;
    uint64_t w0 = w[0];
    if (!s) return w0;
    return (w0 >> s) | (w[1] << (64 - s));
}

// FAST: register-resident core (CLEAN tile, 9 <= w <= 65); otherwise the shared-memory core (any w, any byte)
template <bool CLEAN, bool FAST>
__device__ __forceinline__ void read_tile_body(Tile &t, const ReadSketchArgs &A)
{
    const int tid = threadIdx.x;
    uint16_t *runs = t.pre;                                          // safe: both cores sync before writing runs[]
    uint64_t *run_val = t.canon;                                     // FAST only (canon[] is not used there)
    int halo = -1, n_runs; uint64_t halo_val = 0;
    if (FAST) {
        n_runs = fast_runs<true>(t, runs, run_val, &halo, &halo_val);
    } else {
        phase_canon<CLEAN>(t);
        __syncthreads();
        phase_block_minima<CLEAN>(t);
        __syncthreads();
        n_runs = phase_runs<true, CLEAN>(t, runs, &halo);
    }
    if (n_runs == 0) return;

    if (FAST) { if (tid == 32 * 0 + tile_halo_lanes(t.w)) t.hash[0] = halo >= 0 ? hash_packed_kmer(halo_val, t.k) : 0xFFFFFFFFFFFFFFFFull; }
    else if (tid == 0) t.hash[0] = halo >= 0 ? hash_at<CLEAN>(t, halo) : 0xFFFFFFFFFFFFFFFFull;
    int emitted = 0;
    for (int b0 = 0; b0 < n_runs; b0 += NT) {
        const int j = b0 + tid; const bool have = j < n_runs;
        const int cnt = min(NT, n_runs - b0);
        uint32_t ent = have ? runs[j] : 0;
        uint64_t h = 0;
        if (have) h = FAST ? hash_packed_kmer(run_val[j], t.k) : hash_at<CLEAN>(t, ent & 0x7FFF);
        t.hash[tid + 1] = h;
        __syncthreads();
        uint64_t prev = (ent & 0x8000) ? 0xFFFFFFFFFFFFFFFFull : t.hash[tid];
        uint64_t carry = t.hash[cnt];
        bool emit = have && h != prev;
        if (emit) table_insert(A.table, A.table_mult, A.table_limit, h, A.ctr);
        emitted += __syncthreads_count(emit);
        if (tid == 0) t.hash[0] = carry;
    }
    if (tid == 0 && emitted) atomicAdd(&A.ctr[CTR_READ_EMITTED], (unsigned long long)emitted);
}

__global__ void __launch_bounds__(NT, 3)
read_sketch_kernel(ReadSketchArgs A)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const TileLayout &L = A.layout;
    Tile t = carve(smem, L, A.k, A.w);
    const long long tile = blockIdx.x;
    t.g0 = tile * L.cap - A.w - L.pad;
    t.seq_len = (long long)A.total_bases;
    set_window_bounds(t);
    const int tid = threadIdx.x;

    // ---- read boundaries -> window masks.  A read starting at local base b makes the windows e with b inside their bases
    // (e-w+1, e+k-1], i.e. e in [b-k+1, b+w-2], invalid, and e = b+w-1 the first window of that read.
    const int nwords = (L.M + 31) / 32 + 2;
    for (int i = tid; i < nwords; i += NT) { t.inval[i] = 0; t.firstm[i] = 0; }
    __syncthreads();
    {
        const long long hi = t.g0 + L.NB;
        const int nbits = 32 * nwords;
        uint64_t r0 = A.tile_first_read[tile];
        for (;;) {
            uint64_t r = r0 + tid;
            int past = 1;
            if (r <= A.n_reads) {
                long long off = (long long)A.read_off[r];
                if (off < hi) {
                    past = 0;
                    const int b = (int)(off - t.g0);
                    const int lo = max(b - A.k + 1, 0), hi_b = min(b + A.w - 2, nbits - 1);
                    for (int wi = lo >> 5; wi <= (hi_b >> 5) && lo <= hi_b; ++wi) {
                        uint32_t m = 0xFFFFFFFFu;
                        if (wi == (lo >> 5)) m &= 0xFFFFFFFFu << (lo & 31);
                        if (wi == (hi_b >> 5)) m &= 0xFFFFFFFFu >> (31 - (hi_b & 31));
                        atomicOr(&t.inval[wi], m);
                    }
                    const int f = b + A.w - 1;
                    if (f < nbits) atomicOr(&t.firstm[f >> 5], 1u << (f & 31));
                }
            }
            if (__syncthreads_or(past)) break;
            r0 += NT;
        }
    }
    // ---- stage bases: unaligned 8-byte loads, mask outside [0, total)
    uint32_t dirty_any = 0;
    for (int c = tid; c < L.nchunks; c += NT) {
        long long g = t.g0 + 8ll * c;
        uint64_t v = 0;
        if (g + 8 > 0 && g < t.seq_len) {
            v = load8_unaligned(A.read_bases + g);                   // front padding covers g in [-7, -1]
            if (g < 0) v &= ~0ull << (8 * (int)(-g));
            long long nvalid = t.seq_len - g;                        // bytes [0, nvalid) of the chunk are real
            if (nvalid < 8) v &= (1ull << (8 * nvalid)) - 1;
            v = upcase8(v);
        }
        dirty_any |= stage_chunk(t, c, v);
    }
    // a tile is CLEAN when every staged byte that a valid window can touch is A/C/G/T; padding at either end of
    // the data counts as dirty and sends the (few) boundary tiles through the general path
    if (__syncthreads_or(dirty_any != 0)) read_tile_body<false, false>(t, A);
    else if (tile_fast_w(A.w)) read_tile_body<true, true>(t, A);
    else read_tile_body<true, false>(t, A);
}

// per tile: first read r with read_off[r] >= g0 = tile*cap - w - pad
__global__ void read_tile_dir_kernel(const uint64_t *read_off, uint64_t n_reads, int w, uint64_t n_tiles, uint64_t *tile_first_read)
{
    uint64_t tile = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (tile >= n_tiles) return;
    long long g0 = (long long)tile * tile_cap(w) - w - tile_pad(w);
    uint64_t lo = 0, hi = n_reads + 1;                               // search over read_off[0 .. n_reads]
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if ((long long)read_off[mid] < g0) lo = mid + 1; else hi = mid;
    }
    tile_first_read[tile] = lo;
}

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
            }
        }
        __syncthreads();
    }
    if (tid == 0 && emitted) atomicAdd(&A.chunk_emitted[tr.chunk], (uint32_t)emitted);
}

__global__ void __launch_bounds__(NT, 3)
walk_sketch_kernel(WalkSketchArgs A)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t tile = blockIdx.x;
    const TileRec tr = A.tiles[tile];
    const uint32_t h = tr.walk;
    const long long len = A.walk_len[h];
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
    const uint64_t wbeg = A.walk_off[h], wend = A.walk_off[h + 1];
    const uint64_t s0 = wbeg + tr.first_step;
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
    if (__syncthreads_or(dirty_any != 0)) walk_tile_body<false, false>(t, A, tr, tile);
    else if (tile_fast_w(A.w)) walk_tile_body<true, true>(t, A, tr, tile);
    else walk_tile_body<true, false>(t, A, tr, tile);
}

// ================================================================== hash KAT hook
__global__ void hash_bytes_kernel(const uint8_t *keys, uint64_t n, int len, uint64_t *out)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *p = keys + i * (uint64_t)len;
    uint64_t W[4] = {0, 0, 0, 0};
    bool clean = true;
    for (int j = 0; j < len; ++j) { W[j >> 3] |= (uint64_t)p[j] << (8 * (j & 7)); clean &= is_acgt(p[j]); }
    uint64_t h = murmur3_x64_128_xor(W, len);
    if (clean) {                                                      // cross-check the packed path used by the sketch kernels
        uint64_t km = 0;
        for (int j = 0; j < len; ++j) km = (km << 2) | code2(p[j]);
        uint64_t h2 = hash_packed_kmer(km, len);
        if (h2 != h) h = ~h;                                          // make any divergence visible to the test
    }
    out[i] = h;
}

// ------------------------------------------------------------------ launchers

cudaError_t launch_read_tile_dir(const uint64_t *read_off, uint64_t n_reads, int w, uint64_t n_tiles, uint64_t *out, cudaStream_t st)
{
    if (!n_tiles) return cudaSuccess;
    read_tile_dir_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, st>>>(read_off, n_reads, w, n_tiles, out);
    return cudaGetLastError();
}

cudaError_t launch_read_sketch(const ReadSketchArgs &A, uint64_t n_tiles, cudaStream_t st)
{
    if (!n_tiles) return cudaSuccess;
    size_t smem = (size_t)A.layout.bytes;
    cudaError_t e = cudaFuncSetAttribute(read_sketch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    read_sketch_kernel<<<(unsigned)n_tiles, NT, smem, st>>>(A);
    return cudaGetLastError();
}

cudaError_t launch_walk_sketch(const WalkSketchArgs &A, uint32_t n_tiles, cudaStream_t st)
{
    if (!n_tiles) return cudaSuccess;
    size_t smem = (size_t)A.layout.bytes;
    cudaError_t e = cudaFuncSetAttribute(walk_sketch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    walk_sketch_kernel<<<n_tiles, NT, smem, st>>>(A);
    return cudaGetLastError();
}

cudaError_t launch_hash_bytes(const uint8_t *keys, uint64_t n, int len, uint64_t *out, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    hash_bytes_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(keys, n, len, out);
    return cudaGetLastError();
}

int tile_windows(int w) { return tile_cap(w); }
TileLayout tile_layout(int k, int w, bool walk) { return make_layout(k, w, walk); }

}  // namespace phi
