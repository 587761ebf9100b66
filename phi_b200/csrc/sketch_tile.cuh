// Minimizer tile core shared by the read-sketch and walk-sketch kernels (sm_100a).
//
// A tile owns TILE_W consecutive window end positions e (global k-mer start index of the
// LAST k-mer of the window) of one sequence coordinate system:
//   walk kernel : one haplotype walk  (/root/reference/src/ILP_index.cpp:359-445)
//   read kernel : all reads back to back, windows never cross a read boundary (:447-493)
// Local coordinate p = g - g0 with g0 = tile*TILE_W - w, so the tile sees k-mer positions
// p in [0, M), M = TILE_W + w, and bases p in [0, NB), NB = M + k - 1.  Window end p_e = w-1 is
// the halo window (last window of the previous tile): it is only used to decide whether the
// first window of this tile starts a new run and, if so, what the previous run's hash was.
//
// Reference semantics reproduced here:
//   * canonical k-mer = min(fwd, revcomp) as upper-cased byte strings (:394)
//   * window minimum, rightmost on ties (pop_back on >=, :397)
//   * emit iff hash(min of window i) != hash(min of window i-1); first window of a sequence
//     compares against UINT64_MAX (:383, :413).  prev_hash always equals the hash of the
//     previous window's minimum, so "runs" of equal arg-min position are the unit of work.
//   * non-ACGT bytes take part verbatim ("dirty" k-mers: byte-wise compare + byte-wise hash).
#pragma once
#include "device_common.cuh"

namespace phi {

constexpr int TILE_W = 2048;     // windows per tile
constexpr int NT = 256;          // threads per tile CTA
constexpr int MAX_W = 256;
constexpr int MAX_K = 32;

constexpr uint8_t F_STRAND = 1;  // canonical == reverse complement
constexpr uint8_t F_DIRTY = 2;   // k-mer contains a non-ACGT byte (or padding)

// Dynamic shared memory carve-up.  All offsets in bytes from the (16-aligned) base.
struct TileLayout {
    int M, NB;
    int o_canon, o_pack, o_dirty, o_bnd, o_pre, o_suf, o_arg, o_flag, o_base, o_hash, o_scan, o_stepv, o_steps;
    int bytes;
};

__host__ __device__ inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

__host__ __device__ inline TileLayout make_layout(int k, int w, bool walk)
{
    TileLayout L;
    L.M = TILE_W + w;
    L.NB = L.M + k - 1;
    const int nb8 = align_up(L.NB, 8) + 8;               // bases, padded so chunked stores stay in bounds
    int o = 0;
    auto take = [&o](int bytes) { int at = o; o += align_up(bytes, 16); return at; };   // every section 16-byte aligned
    L.o_canon = take(8 * L.M);
    L.o_hash = take(8 * (NT + 1));
    L.o_pack = take(4 * (nb8 / 16 + 4));
    L.o_dirty = take(4 * (nb8 / 32 + 4));
    L.o_bnd = take(4 * (nb8 / 32 + 4));
    L.o_scan = take(4 * 64);
    L.o_pre = take(2 * align_up(L.M, 2));            // pre and suf stay adjacent: runs[] aliases both
    L.o_suf = take(2 * align_up(L.M, 2));
    L.o_arg = take(2 * align_up(L.M, 2));
    L.o_stepv = take(walk ? 4 * (L.NB + 2) : 0);
    L.o_steps = take(walk ? 2 * (L.NB + 2) : 0);
    L.o_flag = take(L.M);
    L.o_base = take(nb8);
    L.bytes = align_up(o, 16);
    return L;
}

struct Tile {
    // geometry
    int k, w, M, NB;
    long long g0;            // global coordinate of local 0
    long long seq_len;       // WALK: walk length in bases; MULTI: total bases of all reads
    // shared arrays
    uint64_t *canon; uint64_t *hash;
    uint32_t *pack, *dirty, *bnd, *scan;
    uint16_t *pre, *suf, *arg, *steps;
    uint32_t *stepv;
    uint8_t *flag, *base;
};

__device__ __forceinline__ Tile carve(unsigned char *smem, const TileLayout &L, int k, int w)
{
    Tile t;
    t.k = k; t.w = w; t.M = L.M; t.NB = L.NB;
    t.canon = (uint64_t *)(smem + L.o_canon); t.hash = (uint64_t *)(smem + L.o_hash);
    t.pack = (uint32_t *)(smem + L.o_pack); t.dirty = (uint32_t *)(smem + L.o_dirty);
    t.bnd = (uint32_t *)(smem + L.o_bnd); t.scan = (uint32_t *)(smem + L.o_scan);
    t.pre = (uint16_t *)(smem + L.o_pre); t.suf = (uint16_t *)(smem + L.o_suf); t.arg = (uint16_t *)(smem + L.o_arg);
    t.stepv = (uint32_t *)(smem + L.o_stepv); t.steps = (uint16_t *)(smem + L.o_steps);
    t.flag = smem + L.o_flag; t.base = smem + L.o_base;
    return t;
}

// ---- staging helper: 8 upper-cased bytes (as two u32, byte i of the chunk in byte i) for chunk c
// (local bases 8c..8c+7) -> base[], 2-bit pack[], dirty mask.  Padding bytes must be passed as 0.
__device__ __forceinline__ void stage_chunk(const Tile &t, int c, uint32_t lo4, uint32_t hi4)
{
    uint32_t bits = 0, dm = 0;
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t ch = ((i < 4 ? lo4 : hi4) >> (8 * (i & 3))) & 0xFFu;
        bits = (bits << 2) | code2(ch);
        dm |= (is_acgt(ch) ? 0u : 1u) << i;
    }
    ((uint2 *)t.base)[c] = make_uint2(lo4, hi4);
    ((uint16_t *)t.pack)[c ^ 1] = (uint16_t)bits;       // big-endian base order inside each u32
    ((uint8_t *)t.dirty)[c] = (uint8_t)dm;
}

// 2k bits of the packed stream starting at base p, right-aligned
__device__ __forceinline__ uint64_t extract_kmer(const uint32_t *pack, int p, int k)
{
    int i = p >> 4, sh = (p & 15) * 2;
    uint32_t w0 = pack[i], w1 = pack[i + 1], w2 = pack[i + 2];
    uint32_t hi = __funnelshift_l(w1, w0, sh), lo = __funnelshift_l(w2, w1, sh);
    return (((uint64_t)hi << 32) | lo) >> (64 - 2 * k);
}

// any bit set in mask[lo .. hi] (bit positions, inclusive)?
__device__ __forceinline__ bool any_bits(const uint32_t *m, int lo, int hi)
{
    if (hi < lo) return false;
    int wl = lo >> 5, wh = hi >> 5;
    uint32_t first = 0xFFFFFFFFu << (lo & 31), last = 0xFFFFFFFFu >> (31 - (hi & 31));
    if (wl == wh) return (m[wl] & first & last) != 0;
    if (m[wl] & first) return true;
    for (int i = wl + 1; i < wh; ++i) if (m[i]) return true;
    return (m[wh] & last) != 0;
}

// byte-wise three-way compare of the canonical spellings of k-mers a and b (slow path, :394/:397 on raw bytes)
__device__ __noinline__ int cmp_canon_bytes(const Tile &t, int a, int sa, int b, int sb)
{
    for (int j = 0; j < t.k; ++j) {
        uint32_t ca = sa ? comp_byte(t.base[a + t.k - 1 - j]) : t.base[a + j];
        uint32_t cb = sb ? comp_byte(t.base[b + t.k - 1 - j]) : t.base[b + j];
        if (ca != cb) return ca < cb ? -1 : 1;
    }
    return 0;
}

// canon[a] <= canon[b] ?   (va/vb, fa/fb: canon value and flag of a and b)
__device__ __forceinline__ bool canon_le(const Tile &t, int a, uint64_t va, uint32_t fa, int b, uint64_t vb, uint32_t fb)
{
    if (((fa | fb) & F_DIRTY) == 0) return va <= vb;
    return cmp_canon_bytes(t, a, fa & F_STRAND, b, fb & F_STRAND) <= 0;
}
__device__ __forceinline__ bool canon_lt(const Tile &t, int a, uint64_t va, uint32_t fa, int b, uint64_t vb, uint32_t fb)
{
    if (((fa | fb) & F_DIRTY) == 0) return va < vb;
    return cmp_canon_bytes(t, a, fa & F_STRAND, b, fb & F_STRAND) < 0;
}

// hash128_to_64 of the canonical k-mer at local position a
__device__ __noinline__ uint64_t hash_dirty(const Tile &t, int a, int strand)
{
    uint64_t W[4] = {0, 0, 0, 0};
    for (int j = 0; j < t.k; ++j) {
        uint64_t c = strand ? comp_byte(t.base[a + t.k - 1 - j]) : t.base[a + j];
        W[j >> 3] |= c << (8 * (j & 7));
    }
    return murmur3_x64_128_xor(W, t.k);
}
__device__ __forceinline__ uint64_t hash_at(const Tile &t, int a)
{
    uint32_t f = t.flag[a];
    if (f & F_DIRTY) return hash_dirty(t, a, f & F_STRAND);
    return hash_packed_kmer(t.canon[a], t.k);
}

// ---- phase 3: canonical k-mers for p in [0, M)
__device__ __forceinline__ void phase_canon(const Tile &t)
{
    for (int p = threadIdx.x; p < t.M; p += NT) {
        uint64_t fwd = extract_kmer(t.pack, p, t.k);
        uint32_t d = __funnelshift_r(t.dirty[p >> 5], t.dirty[(p >> 5) + 1], p & 31);
        if (t.k < 32) d &= (1u << t.k) - 1;
        uint64_t cv; uint32_t f;
        if (d == 0) {
            uint64_t rc = revcomp2(fwd, t.k);
            f = rc < fwd ? F_STRAND : 0;
            cv = rc < fwd ? rc : fwd;
        } else {
            // std::min(fwd, rev): rev only if strictly smaller (:394)
            int c = cmp_canon_bytes(t, p, 1, p, 0);
            f = F_DIRTY | (c < 0 ? F_STRAND : 0);
            cv = 0;
        }
        t.canon[p] = cv; t.flag[p] = (uint8_t)f;
    }
}

// ---- phase 4: van Herk / Gil-Werman block prefix (rightmost-min) and suffix (rightmost-min) arg-minima
__device__ __forceinline__ void phase_block_minima(const Tile &t)
{
    const int nblk = (t.M + t.w - 1) / t.w;
    for (int id = threadIdx.x; id < 2 * nblk; id += NT) {
        if (id < nblk) {                                     // prefix: later position wins ties
            int b0 = id * t.w, b1 = min(b0 + t.w, t.M);
            int cur = b0; uint64_t cv = t.canon[b0]; uint32_t cf = t.flag[b0];
            t.pre[b0] = (uint16_t)b0;
            for (int p = b0 + 1; p < b1; ++p) {
                uint64_t v = t.canon[p]; uint32_t f = t.flag[p];
                if (canon_le(t, p, v, f, cur, cv, cf)) { cur = p; cv = v; cf = f; }
                t.pre[p] = (uint16_t)cur;
            }
        } else {                                             // suffix: earlier position wins only if strictly smaller
            int b0 = (id - nblk) * t.w, b1 = min(b0 + t.w, t.M);
            int cur = b1 - 1; uint64_t cv = t.canon[cur]; uint32_t cf = t.flag[cur];
            t.suf[cur] = (uint16_t)cur;
            for (int p = b1 - 2; p >= b0; --p) {
                uint64_t v = t.canon[p]; uint32_t f = t.flag[p];
                if (canon_lt(t, p, v, f, cur, cv, cf)) { cur = p; cv = v; cf = f; }
                t.suf[p] = (uint16_t)cur;
            }
        }
    }
}

// ---- phase 5: arg-min (rightmost) of every window end p_e in [w-1, M)
__device__ __forceinline__ void phase_window_argmin(const Tile &t)
{
    for (int e = t.w - 1 + threadIdx.x; e < t.M; e += NT) {
        int a = t.suf[e - t.w + 1], b = t.pre[e];
        int r = b;
        if (a != b) r = canon_le(t, b, t.canon[b], t.flag[b], a, t.canon[a], t.flag[a]) ? b : a;
        t.arg[e] = (uint16_t)r;
    }
}

// Sequence model: which windows exist, and which is the first of its sequence.
template <bool MULTI>
struct SeqModel {
    // k-mer positions of window e are [e-w+1, e]; bases [e-w+1, e+k-1]
    __device__ static __forceinline__ bool window_valid(const Tile &t, int e)
    {
        long long gs = t.g0 + e - t.w + 1, ge = t.g0 + e + t.k;           // [gs, ge) bases
        if (gs < 0 || ge > t.seq_len) return false;
        if (MULTI) return !any_bits(t.bnd, e - t.w + 2, e + t.k - 1);     // no read starts strictly inside
        return true;
    }
    __device__ static __forceinline__ bool window_first(const Tile &t, int e)
    {
        if (MULTI) { int s = e - t.w + 1; return (t.bnd[s >> 5] >> (s & 31)) & 1u; }
        return t.g0 + e == t.w - 1;
    }
};

// ---- phase 6: run starts among the tile's own windows e in [w, M), compacted in position order.
// Run entry: arg-min position | 0x8000 if the run starts at the first window of its sequence.
// Returns the number of runs (block-uniform).  runs[] aliases pre[]/suf[] (dead by now).
template <bool MULTI>
__device__ __forceinline__ int phase_runs(const Tile &t, uint16_t *runs)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int PER_WARP = TILE_W / (NT / 32);                          // 256 windows per warp
    constexpr int ROUNDS = PER_WARP / 32;                                 // 8
    uint32_t ballots[ROUNDS];
    uint16_t mine[ROUNDS];
    int total = 0;
    #pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        int e = t.w + wid * PER_WARP + r * 32 + lane;
        bool start = false; uint16_t ent = 0;
        if (SeqModel<MULTI>::window_valid(t, e)) {
            int a = t.arg[e];
            if (SeqModel<MULTI>::window_first(t, e)) { start = true; ent = (uint16_t)(a | 0x8000); }
            else if (a != t.arg[e - 1]) { start = true; ent = (uint16_t)a; }
        }
        ballots[r] = __ballot_sync(0xFFFFFFFFu, start);
        mine[r] = ent;
        total += __popc(ballots[r]);
    }
    if (lane == 0) t.scan[wid] = total;
    __syncthreads();                                                      // also: arg[] reads done before runs[] (aliasing pre/suf only)
    int off = 0, all = 0;
    #pragma unroll
    for (int i = 0; i < NT / 32; ++i) { int c = t.scan[i]; if (i < wid) off += c; all += c; }
    #pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        if ((ballots[r] >> lane) & 1u) runs[off + __popc(ballots[r] & lanemask_lt())] = mine[r];
        off += __popc(ballots[r]);
    }
    __syncthreads();
    return all;
}

// Hash of the run preceding the tile's first window (the halo window's minimum), or UINT64_MAX when the
// tile starts its sequence coordinate (no previous window exists).  Only meaningful if the first run of the
// tile is not flagged "first of sequence" — flagged runs ignore it.
template <bool MULTI>
__device__ __forceinline__ uint64_t halo_prev_hash(const Tile &t)
{
    int e = t.w - 1;
    if (!SeqModel<MULTI>::window_valid(t, e)) return 0xFFFFFFFFFFFFFFFFull;
    return hash_at(t, t.arg[e]);
}

}  // namespace phi
