// Minimizer tile core shared by the read-sketch and walk-sketch kernels (sm_100a) — v2.
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
//
// v2 (instruction diet, see profiles/r1_walk_kernel_phases_v1.txt): the layout comes from the
// host; tiles without any non-ACGT byte run a CLEAN specialisation with no flag traffic; canonical
// k-mers are rolled 8 per thread and stored with 4 STS.128 into a padded, conflict-free array;
// window arg-min and run detection are fused (warp shuffles instead of an arg[] array).
#pragma once
#include "device_common.cuh"
#include "kernels.h"

#ifndef PHI_TILE_THREADS
#error "define PHI_TILE_THREADS (threads of the tile CTA) before including sketch_tile.cuh"
#endif

namespace phi {
namespace {   // every translation unit gets its own copy, specialised for its tile size

constexpr int NT = PHI_TILE_THREADS;                   // threads per tile CTA
constexpr int TILE_W = TILE_WINDOWS_PER_THREAD * NT;   // windows per tile on the general path
constexpr int MAX_W = 256;
constexpr int MAX_K = 255;                             // k-mers of up to 32 bases are packed 2-bit; longer ones are compared and hashed byte-wise
constexpr int MAX_PACKED_K = 32;

constexpr uint8_t F_STRAND = 1;  // canonical == reverse complement
constexpr uint8_t F_DIRTY = 2;   // k-mer contains a non-ACGT byte (or padding)

// Dynamic shared memory carve-up: struct TileLayout lives in kernels.h (computed once on the host, passed by value).
static inline int align_up_h(int x, int a) { return (x + a - 1) / a * a; }

static inline TileLayout make_layout(int k, int w, bool walk)
{
    TileLayout L;
    L.pad = tile_pad(w); L.cap = tile_cap(w, NT);
    L.M = L.cap + w + L.pad;
    L.M8 = align_up_h(L.M, 8);
    L.NB = L.M + k - 1;
    L.nchunks = (L.NB + 7) / 8;
    const int nb8 = 8 * L.nchunks + 8;                   // bases, padded so chunked stores stay in bounds
    int o = 0;
    auto take = [&o](int bytes) { int at = o; o += align_up_h(bytes, 16); return at; };   // every section 16-byte aligned
    L.o_canon = take(8 * (L.M8 + 2 * (L.M8 / 8)));       // padded: idx(p) = p + 2*(p>>3)
    L.o_hash = take(8 * (NT + 1));
    L.o_pack = take(4 * (nb8 / 16 + 12));                // slack: the register core reads up to 8 lanes x 8 positions past the tile's last k-mer
    L.o_dirty = take(4 * (nb8 / 32 + 4));
    L.o_bnd = take(4 * (nb8 / 32 + 8));
    L.o_first = take(4 * (nb8 / 32 + 8));
    L.o_scan = take(4 * 96);
    L.o_pre = take(2 * (L.M8 + 8));                      // runs[] aliases pre (TILE_W + 1 entries <= M)
    L.o_suf = take(2 * (L.M8 + 8));
    L.o_flag = take(L.M8 + 8);
    L.o_base = take(nb8);
    L.o_stepv = take(walk ? 4 * (L.NB + 2) : 0);
    L.o_steps = take(walk ? 2 * (L.NB + 2) : 0);
    L.o_cfirst = take(walk ? 2 * (L.nchunks + 2) : 0);
    L.o_cmask = take(walk ? (L.nchunks + 2) : 0);
    L.o_raw = take(walk ? 0 : 8 * L.nchunks + 64);
    L.bytes = align_up_h(o, 16);
    return L;
}

struct Tile {
    // geometry
    int k, w, M, M8, NB;
    int pad, e_halo, e_own0; // front padding; local index of the halo window (w-1+pad) and of the first own window (w+pad)
    long long g0;            // global coordinate of local 0
    long long seq_len;       // WALK: walk length in bases; MULTI: total bases of all reads
    int e_lo, e_hi;          // WALK: local window ends e with a valid window are [e_lo, e_hi); e_first: the sequence's first window
    int e_first;
    // shared arrays
    uint64_t *canon; uint64_t *hash;
    uint32_t *pack, *dirty, *scan;
    uint32_t *inval, *firstm;   // MULTI: bit e set iff window e spans a read boundary / is the first window of its read
    uint16_t *pre, *suf, *steps, *cfirst;
    uint32_t *stepv;
    uint8_t *flag, *base, *cmask;
};

__device__ __forceinline__ Tile carve(unsigned char *smem, const TileLayout &L, int k, int w)
{
    Tile t;
    t.k = k; t.w = w; t.M = L.M; t.M8 = L.M8; t.NB = L.NB;
    t.pad = L.pad; t.e_halo = w - 1 + L.pad; t.e_own0 = w + L.pad;
    t.canon = (uint64_t *)(smem + L.o_canon); t.hash = (uint64_t *)(smem + L.o_hash);
    t.pack = (uint32_t *)(smem + L.o_pack); t.dirty = (uint32_t *)(smem + L.o_dirty);
    t.inval = (uint32_t *)(smem + L.o_bnd); t.firstm = (uint32_t *)(smem + L.o_first); t.scan = (uint32_t *)(smem + L.o_scan);
    t.pre = (uint16_t *)(smem + L.o_pre); t.suf = (uint16_t *)(smem + L.o_suf);
    t.stepv = (uint32_t *)(smem + L.o_stepv); t.steps = (uint16_t *)(smem + L.o_steps);
    t.cfirst = (uint16_t *)(smem + L.o_cfirst); t.cmask = smem + L.o_cmask;
    t.flag = smem + L.o_flag; t.base = smem + L.o_base;
    return t;
}

// padded canon index: 10 slots per 8 positions -> 80-byte thread stride, conflict-free STS.128 / LDS.64
__device__ __forceinline__ int cidx(int p) { return p + 2 * (p >> 3); }

// upper-case 8 packed bytes (SWAR; bytes >= 0x80 untouched, like ::toupper in the C locale)
__device__ __forceinline__ uint64_t upcase8(uint64_t x)
{
    const uint64_t lo7 = 0x7F7F7F7F7F7F7F7Full;
    uint64_t t = x & lo7;
    uint64_t ge_a = t + 0x1F1F1F1F1F1F1F1Full;          // bit7 set iff byte >= 'a' (0x61)
    uint64_t gt_z = t + 0x0505050505050505ull;          // bit7 set iff byte >  'z' (0x7A)
    uint64_t lower = ge_a & ~gt_z & ~x & 0x8080808080808080ull;
    return x ^ (lower >> 2);                            // clear 0x20
}

// ---- staging helper: 8 upper-cased bytes for chunk c (local bases 8c..8c+7, byte i of v = base i)
// -> base[], 2-bit pack[], dirty mask.  Padding bytes must be passed as 0.  Returns the chunk's dirty mask.
__device__ __forceinline__ uint32_t stage_chunk(const Tile &t, int c, uint64_t v)
{
    const uint32_t lo4 = (uint32_t)v, hi4 = (uint32_t)(v >> 32);
    // codes: ((c>>1)^(c>>2))&3 per byte, then gather 4 codes into one byte with a multiply (no carries: 2-bit fields)
    uint32_t cl = ((lo4 >> 1) ^ (lo4 >> 2)) & 0x03030303u, ch = ((hi4 >> 1) ^ (hi4 >> 2)) & 0x03030303u;
    uint32_t bits = (((cl * 0x40100401u) >> 24) << 8) | ((ch * 0x40100401u) >> 24);
    uint32_t dm = 0;
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t b = ((i < 4 ? lo4 : hi4) >> (8 * (i & 3))) & 0xFFu;
        dm |= (is_acgt(b) ? 0u : 1u) << i;
    }
    ((uint2 *)t.base)[c] = make_uint2(lo4, hi4);
    ((uint16_t *)t.pack)[c ^ 1] = (uint16_t)bits;       // big-endian base order inside each u32
    ((uint8_t *)t.dirty)[c] = (uint8_t)dm;
    // bytes at or beyond NB are slack of the last chunk: no k-mer of the tile can touch them, so they must not make the tile dirty
    const int live = t.NB - 8 * c;
    return live >= 8 ? dm : (dm & ((1u << (live > 0 ? live : 0)) - 1u));
}

// 2k bits of the packed stream starting at base p, right-aligned
__device__ __forceinline__ uint64_t extract_kmer(const uint32_t *pack, int p, int k)
{
    int i = p >> 4, sh = (p & 15) * 2;
    uint32_t w0 = pack[i], w1 = pack[i + 1], w2 = pack[i + 2];
    uint32_t hi = __funnelshift_l(w1, w0, sh), lo = __funnelshift_l(w2, w1, sh);
    return (((uint64_t)hi << 32) | lo) >> (64 - 2 * k);
}
// the 8 bases p..p+7 as 16 bits (base p on top)
__device__ __forceinline__ uint32_t extract8(const uint32_t *pack, int p)
{
    int i = p >> 4, sh = (p & 15) * 2;
    return __funnelshift_l(pack[i + 1], pack[i], sh) >> 16;
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
template <bool CLEAN>
__device__ __forceinline__ bool canon_le(const Tile &t, int a, uint64_t va, uint32_t fa, int b, uint64_t vb, uint32_t fb)
{
    if (CLEAN || ((fa | fb) & F_DIRTY) == 0) return va <= vb;
    return cmp_canon_bytes(t, a, fa & F_STRAND, b, fb & F_STRAND) <= 0;
}
template <bool CLEAN>
__device__ __forceinline__ bool canon_lt(const Tile &t, int a, uint64_t va, uint32_t fa, int b, uint64_t vb, uint32_t fb)
{
    if (CLEAN || ((fa | fb) & F_DIRTY) == 0) return va < vb;
    return cmp_canon_bytes(t, a, fa & F_STRAND, b, fb & F_STRAND) < 0;
}

// hash128_to_64 of the canonical k-mer at local position a
__device__ __noinline__ uint64_t hash_dirty(const Tile &t, int a, int strand)
{
    const uint8_t *base = t.base; const int k = t.k;
    return murmur3_x64_128_xor_bytes([&](int j) -> uint32_t { return strand ? comp_byte(base[a + k - 1 - j]) : (uint32_t)base[a + j]; }, k);
}
template <bool CLEAN>
__device__ __forceinline__ uint64_t hash_at(const Tile &t, int a)
{
    if (!CLEAN) {
        uint32_t f = t.flag[a];
        if (f & F_DIRTY) return hash_dirty(t, a, f & F_STRAND);
    }
    return hash_packed_kmer(t.canon[cidx(a)], t.k);
}

// ---- phase: canonical k-mers for p in [0, M8), 8 consecutive positions per thread, rolled
template <bool CLEAN>
__device__ __forceinline__ void phase_canon(const Tile &t)
{
    const int k = t.k;
    const uint64_t kmask = k == 32 ? ~0ull : (1ull << (2 * k)) - 1;
    const int top = 2 * (k - 1);
    for (int g8 = threadIdx.x; g8 < t.M8 / 8; g8 += NT) {
        const int p0 = 8 * g8;
        if (!CLEAN && k > MAX_PACKED_K) {
            // long k-mers: every one is spelled out (the tile is treated like one full of non-ACGT bytes: byte-wise compare and hash)
            uint64_t fl = 0;
            #pragma unroll 1
            for (int i = 0; i < 8; ++i) {
                const int c = cmp_canon_bytes(t, p0 + i, 1, p0 + i, 0);        // std::min(fwd, rev): rev only if strictly smaller (:394)
                fl |= (uint64_t)(F_DIRTY | (c < 0 ? F_STRAND : 0)) << (8 * i);
            }
            *(uint64_t *)(t.flag + p0) = fl;
            ulonglong2 *dst = (ulonglong2 *)(t.canon + cidx(p0));
            dst[0] = dst[1] = dst[2] = dst[3] = make_ulonglong2(0ull, 0ull);
            continue;
        }
        uint64_t fwd = extract_kmer(t.pack, p0, k);
        uint64_t rc = revcomp2(fwd, k);
        const uint32_t nxt = extract8(t.pack, p0 + k);                 // bases entering at steps 1..7 (+1 spare)
        uint64_t cv[8];
        cv[0] = rc < fwd ? rc : fwd;
        uint32_t strand = rc < fwd ? 1u : 0u;
        #pragma unroll
        for (int i = 1; i < 8; ++i) {
            uint64_t code = (nxt >> (16 - 2 * i)) & 3u;
            fwd = ((fwd << 2) | code) & kmask;
            rc = (rc >> 2) | ((3ull - code) << top);
            cv[i] = rc < fwd ? rc : fwd;
            strand |= (rc < fwd ? 1u : 0u) << i;
        }
        if (!CLEAN) {
            // non-ACGT bytes: k-mer i is dirty iff any of dirty bits [p0+i, p0+i+k) is set
            const int wi = p0 >> 5, sh = p0 & 31;
            uint64_t d64 = ((uint64_t)__funnelshift_r(t.dirty[wi + 1], t.dirty[wi + 2], sh) << 32) | __funnelshift_r(t.dirty[wi], t.dirty[wi + 1], sh);
            const uint64_t km = k == 32 ? 0xFFFFFFFFull : (1ull << k) - 1;
            uint64_t fl = 0;
            #pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t f = (strand >> i) & 1u;
                if ((d64 >> i) & km) {
                    int c = cmp_canon_bytes(t, p0 + i, 1, p0 + i, 0);  // std::min(fwd, rev): rev only if strictly smaller (:394)
                    f = F_DIRTY | (c < 0 ? F_STRAND : 0);
                    cv[i] = 0;
                }
                fl |= (uint64_t)f << (8 * i);
            }
            *(uint64_t *)(t.flag + p0) = fl;
        }
        ulonglong2 *dst = (ulonglong2 *)(t.canon + cidx(p0));
        dst[0] = make_ulonglong2(cv[0], cv[1]); dst[1] = make_ulonglong2(cv[2], cv[3]);
        dst[2] = make_ulonglong2(cv[4], cv[5]); dst[3] = make_ulonglong2(cv[6], cv[7]);
    }
}

// ---- phase: van Herk / Gil-Werman block prefix (rightmost-min) and suffix (rightmost-min) arg-minima
template <bool CLEAN>
__device__ __forceinline__ void phase_block_minima(const Tile &t)
{
    const int nblk = (t.M + t.w - 1) / t.w;
    for (int id = threadIdx.x; id < 2 * nblk; id += NT) {
        if (id < nblk) {                                     // prefix: later position wins ties
            int b0 = id * t.w, b1 = min(b0 + t.w, t.M);
            int cur = b0; uint64_t cv = t.canon[cidx(b0)]; uint32_t cf = CLEAN ? 0 : t.flag[b0];
            t.pre[b0] = (uint16_t)b0;
            for (int p = b0 + 1; p < b1; ++p) {
                uint64_t v = t.canon[cidx(p)]; uint32_t f = CLEAN ? 0 : t.flag[p];
                if (canon_le<CLEAN>(t, p, v, f, cur, cv, cf)) { cur = p; cv = v; cf = f; }
                t.pre[p] = (uint16_t)cur;
            }
        } else {                                             // suffix: earlier position wins only if strictly smaller
            int b0 = (id - nblk) * t.w, b1 = min(b0 + t.w, t.M);
            int cur = b1 - 1; uint64_t cv = t.canon[cidx(cur)]; uint32_t cf = CLEAN ? 0 : t.flag[cur];
            t.suf[cur] = (uint16_t)cur;
            for (int p = b1 - 2; p >= b0; --p) {
                uint64_t v = t.canon[cidx(p)]; uint32_t f = CLEAN ? 0 : t.flag[p];
                if (canon_lt<CLEAN>(t, p, v, f, cur, cv, cf)) { cur = p; cv = v; cf = f; }
                t.suf[p] = (uint16_t)cur;
            }
        }
    }
}

// arg-min (rightmost) of the window ending at e
template <bool CLEAN>
__device__ __forceinline__ int window_argmin(const Tile &t, int e)
{
    int a = t.suf[e - t.w + 1], b = t.pre[e];
    if (a == b) return b;
    uint32_t fa = CLEAN ? 0 : t.flag[a], fb = CLEAN ? 0 : t.flag[b];
    return canon_le<CLEAN>(t, b, t.canon[cidx(b)], fb, a, t.canon[cidx(a)], fa) ? b : a;
}

// Sequence model: which windows exist, and which is the first of its sequence.
template <bool MULTI>
struct SeqModel {
    // k-mer positions of window e are [e-w+1, e]; bases [e-w+1, e+k-1]
    __device__ static __forceinline__ bool window_valid(const Tile &t, int e)
    {
        if (e < t.e_lo || e >= t.e_hi) return false;
        if (MULTI) return !((t.inval[e >> 5] >> (e & 31)) & 1u);          // no read starts strictly inside
        return true;
    }
    __device__ static __forceinline__ bool window_first(const Tile &t, int e)
    {
        if (MULTI) return (t.firstm[e >> 5] >> (e & 31)) & 1u;
        return e == t.e_first;
    }
};

// window-validity bounds in local coordinates from g0 / seq_len (call once per tile)
__device__ __forceinline__ void set_window_bounds(Tile &t)
{
    // valid: g0 + e - w + 1 >= 0  and  g0 + e + k <= seq_len
    long long lo = (long long)t.w - 1 - t.g0, hi = t.seq_len - t.k - t.g0 + 1;
    t.e_lo = (int)max(lo, (long long)t.e_halo);
    t.e_hi = (int)min(hi, (long long)t.M);
    t.e_first = (int)min(max(lo, -1ll), (long long)t.M + 1);               // only meaningful for the walk model
}

// ---- phase: window arg-minima + run starts among the tile's own windows e in [w, M), compacted in position order.
// Run entry: arg-min position | 0x8000 if the run starts at the first window of its sequence.
// Returns the number of runs (block-uniform).  runs[] aliases pre[]: the caller passes a separate array.
// *halo_arg receives the arg-min of the halo window (e = w-1) if that window is valid, else -1 (thread 0 only).
template <bool MULTI, bool CLEAN>
__device__ __forceinline__ int phase_runs(const Tile &t, uint16_t *runs, int *halo_arg)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int PER_WARP = TILE_W / (NT / 32);                          // 256 windows per warp
    constexpr int ROUNDS = PER_WARP / 32;                                 // 8
    uint32_t ballots[ROUNDS];
    uint16_t mine[ROUNDS];
    int total = 0;
    int carry = -1;                                                       // arg-min of the window before this lane-0's window
    {
        int e0 = t.e_own0 + wid * PER_WARP - 1;                           // window preceding the warp's first
        if (lane == 0 && SeqModel<MULTI>::window_valid(t, e0)) carry = window_argmin<CLEAN>(t, e0);
        if (threadIdx.x == 0) *halo_arg = carry;
    }
    #pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int e = t.e_own0 + wid * PER_WARP + r * 32 + lane;
        if (e - lane >= t.e_hi) { ballots[r] = 0; mine[r] = 0; continue; }  // warp-uniform: a short tile has no windows here (nor further on)
        const bool valid = SeqModel<MULTI>::window_valid(t, e);
        int a = valid ? window_argmin<CLEAN>(t, e) : -1;
        int prev = __shfl_up_sync(0xFFFFFFFFu, a, 1);
        if (lane == 0) prev = carry;
        carry = __shfl_sync(0xFFFFFFFFu, a, 31);
        bool start = false; uint16_t ent = 0;
        if (valid) {
            if (SeqModel<MULTI>::window_first(t, e)) { start = true; ent = (uint16_t)(a | 0x8000); }
            else if (a != prev) { start = true; ent = (uint16_t)a; }
        }
        ballots[r] = __ballot_sync(0xFFFFFFFFu, start);
        mine[r] = ent;
        total += __popc(ballots[r]);
    }
    if (lane == 0) t.scan[wid] = total;
    __syncthreads();                                                      // all pre[]/suf[] reads are done: runs[] may alias pre[]
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

// =====================================================================================================
// Register-resident core (CLEAN tiles, 9 <= w <= 65): canonical k-mers, window arg-minima and run starts
// without the canon[] / pre[] / suf[] round trips through shared memory.
//
// Lane l of warp wid owns the 8 consecutive k-mer positions  p = wid*UW + 8*l + i  (UW = 8*(32-HL) windows
// per warp, HL = ceil((w-1)/8) halo lanes).  The window ending at position e = 8*l + i of a warp starts at
// 8*(l-q) + (i-r) with w-1 = 8q + r, so its minimum is  suffix-min of one earlier lane (from offset i-r, or
// i-r+8 one lane further back)  (+)  the full-lane minima in between  (+)  the prefix-min of the own lane up
// to i, combined left to right with the right operand winning ties (rightmost minimal k-mer, :397).  Suffix
// and full-lane minima travel by warp shuffles; the register index i-r is made a compile-time constant by
// dispatching on r.
// =====================================================================================================
struct MinP { uint64_t v; int pos; };
__device__ __forceinline__ MinP comb(MinP left, MinP right) { return right.v <= left.v ? right : left; }   // right wins ties

template <int R>
__device__ __forceinline__ void window_minima(const uint64_t (&pv)[8], uint32_t pidx, const uint64_t (&sv)[8], uint32_t sidx, int q, int lane_pos0,
                                              uint64_t (&wv)[8], int (&wp)[8])
{
    const unsigned FULL = 0xFFFFFFFFu;
    // full-lane minima of lanes l-1 .. l-q, folded left to right: FA = lanes l-q+1 .. l-1, FB = lanes l-q .. l-1
    MinP FA; FA.v = 0xFFFFFFFFFFFFFFFFull; FA.pos = -1;
    MinP FB = FA;
    const uint64_t fullv = pv[7]; const int fullp = lane_pos0 + (int)(pidx >> 21);
    for (int j = q; j >= 1; --j) {
        MinP F; F.v = __shfl_up_sync(FULL, fullv, j); F.pos = __shfl_up_sync(FULL, fullp, j);
        if (j == q) FB = F; else { FA = comb(FA, F); FB = comb(FB, F); }
    }
    const uint32_t sidx_a = __shfl_up_sync(FULL, sidx, q), sidx_b = __shfl_up_sync(FULL, sidx, q + 1);
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
        MinP S, cur;
        if (i >= R) {                                        // suffix of lane l-q from offset i-R, then lanes l-q+1 .. l-1
            S.v = __shfl_up_sync(FULL, sv[(i - R) & 7], q);
            S.pos = lane_pos0 - 8 * q + (int)((sidx_a >> (3 * ((i - R) & 7))) & 7u);
            cur = comb(S, FA);
        } else {                                             // suffix of lane l-q-1 from offset i-R+8, then lanes l-q .. l-1
            S.v = __shfl_up_sync(FULL, sv[(i - R + 8) & 7], q + 1);
            S.pos = lane_pos0 - 8 * (q + 1) + (int)((sidx_b >> (3 * ((i - R + 8) & 7))) & 7u);
            cur = comb(S, FB);
        }
        MinP P; P.v = pv[i]; P.pos = lane_pos0 + (int)((pidx >> (3 * i)) & 7u);
        cur = comb(cur, P);
        wv[i] = cur.v; wp[i] = cur.pos;
    }
}

// Runs of the tile's own windows, compacted in position order: runs[j] = arg-min position (| 0x8000: first window of its
// sequence), run_val[j] = the canonical k-mer there.  *halo_pos / *halo_val: arg-min of the halo window (-1 if it does not exist).
// Returns the number of runs (block-uniform).
template <bool MULTI>
__device__ __forceinline__ int fast_runs(const Tile &t, uint16_t *runs, uint64_t *run_val, int *halo_pos, uint64_t *halo_val)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int k = t.k, w = t.w;
    const int HL = tile_halo_lanes(w), UW = 8 * (32 - HL);
    const int q = (w - 1) >> 3, r = (w - 1) & 7;
    const int p0 = wid * UW + 8 * lane;                       // local position of this lane's first k-mer
    const bool warp_live = wid * UW + 8 * HL < t.e_hi;        // warp-uniform: the warp has at least one window below e_hi

    uint64_t wv[8]; int wp[8];
    if (warp_live) {
        // ---- canonical k-mers of positions p0 .. p0+7, rolled
        uint64_t cv[8];
        {
            const uint64_t kmask = k == 32 ? ~0ull : (1ull << (2 * k)) - 1;
            const int top = 2 * (k - 1);
            uint64_t fwd = extract_kmer(t.pack, p0, k);
            uint64_t rc = revcomp2(fwd, k);
            const uint32_t nxt = extract8(t.pack, p0 + k);
            cv[0] = rc < fwd ? rc : fwd;
            #pragma unroll
            for (int i = 1; i < 8; ++i) {
                uint64_t code = (nxt >> (16 - 2 * i)) & 3u;
                fwd = ((fwd << 2) | code) & kmask;
                rc = (rc >> 2) | ((3ull - code) << top);
                cv[i] = rc < fwd ? rc : fwd;
            }
        }
        // ---- prefix (later position wins ties) and suffix (earlier position wins only if strictly smaller) minima inside the lane
        uint64_t pv[8], sv[8]; uint32_t pidx = 0, sidx = 7u << 21;
        pv[0] = cv[0]; sv[7] = cv[7];
        {
            uint32_t pi = 0, si = 7;
            #pragma unroll
            for (int i = 1; i < 8; ++i) {
                bool take = cv[i] <= pv[i - 1];
                pv[i] = take ? cv[i] : pv[i - 1]; pi = take ? (uint32_t)i : pi; pidx |= pi << (3 * i);
            }
            #pragma unroll
            for (int i = 6; i >= 0; --i) {
                bool take = cv[i] < sv[i + 1];
                sv[i] = take ? cv[i] : sv[i + 1]; si = take ? (uint32_t)i : si; sidx |= si << (3 * i);
            }
        }
        switch (r) {
            case 0: window_minima<0>(pv, pidx, sv, sidx, q, p0, wv, wp); break;
            case 1: window_minima<1>(pv, pidx, sv, sidx, q, p0, wv, wp); break;
            case 2: window_minima<2>(pv, pidx, sv, sidx, q, p0, wv, wp); break;
            case 3: window_minima<3>(pv, pidx, sv, sidx, q, p0, wv, wp); break;
            case 4: window_minima<4>(pv, pidx, sv, sidx, q, p0, wv, wp); break;
            case 5: window_minima<5>(pv, pidx, sv, sidx, q, p0, wv, wp); break;
            case 6: window_minima<6>(pv, pidx, sv, sidx, q, p0, wv, wp); break;
            default: window_minima<7>(pv, pidx, sv, sidx, q, p0, wv, wp); break;
        }
    } else {
        #pragma unroll
        for (int i = 0; i < 8; ++i) { wp[i] = -1; wv[i] = 0; }
    }
    // windows that exist: inside the sequence (and, MULTI, not across a read boundary); none in the halo lanes.  One byte per lane.
    uint32_t valid8 = 0, first8 = 0;
    if (warp_live && lane >= HL) {
        const int lo_i = min(max(t.e_lo - p0, 0), 8), hi_i = min(max(t.e_hi - p0, 0), 8);
        valid8 = ((1u << hi_i) - 1u) & ~((1u << lo_i) - 1u);
        if (MULTI) { valid8 &= ~(uint32_t)((const uint8_t *)t.inval)[p0 >> 3]; first8 = ((const uint8_t *)t.firstm)[p0 >> 3]; }
        else { const int f = t.e_first - p0; first8 = (f >= 0 && f < 8) ? 1u << f : 0u; }
    }
    #pragma unroll
    for (int i = 0; i < 8; ++i) if (!((valid8 >> i) & 1u)) wp[i] = -1;
    // the window before a warp's first one is the last window of the previous warp
    int *warp_last = (int *)(t.scan + 64);
    if (lane == 31) warp_last[wid] = wp[7];
    __syncthreads();
    int prev = __shfl_up_sync(FULL, wp[7], 1);
    if (lane == HL) prev = wid ? warp_last[wid - 1] : -1;
    uint32_t changed = 0;                                       // bit i: arg-min of window i differs from the previous window's
    #pragma unroll
    for (int i = 0; i < 8; ++i) { changed |= (wp[i] != prev ? 1u : 0u) << i; prev = wp[i]; }
    const uint32_t firsts = first8 & valid8;
    uint32_t starts = valid8 & (firsts | changed);
    if (wid == 0 && lane == HL) starts &= ~1u;                  // the halo window only hands its arg-min on
    if (wid == 0 && lane == HL) { *halo_pos = wp[0]; *halo_val = wv[0]; }      // window e_halo = lane HL, i = 0 of warp 0
    // ---- compaction in position order: lanes in order, inside a lane by i
    const int cnt = __popc(starts);
    int inc = cnt;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int tt = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += tt; }
    if (lane == 31) t.scan[wid] = inc;
    __syncthreads();
    int off = inc - cnt, all = 0;
    #pragma unroll
    for (int j = 0; j < NT / 32; ++j) { int c = t.scan[j]; if (j < wid) off += c; all += c; }
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
        if ((starts >> i) & 1u) {
            runs[off] = (uint16_t)(wp[i] | (((firsts >> i) & 1u) ? 0x8000 : 0));
            run_val[off] = wv[i];
            ++off;
        }
    }
    __syncthreads();
    return all;
}

}  // namespace
}  // namespace phi
