// Hand-written device primitives used around the sketch kernels: exclusive scans, a stable LSD radix
// sort on u64 keys (optionally carrying a u32 value), compaction of the spectrum table, and the radix
// directory that turns the sorted spectrum into an O(1)-expected rank lookup.
// All of this is HBM-bound integer work: coalesced streaming loads/stores, shared-memory histograms,
// warp match/ballot ranking; no tensor cores (nothing here is a contraction).
#include "kernels.h"
#include "device_common.cuh"

namespace phi {

#define PHI_LAUNCH_CHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return e_; if (launches) ++*launches; } while (0)

// ------------------------------------------------------------------ scan
constexpr int SCAN_T = 256, SCAN_I = 8, SCAN_B = SCAN_T * SCAN_I;

template <class T>
__device__ __forceinline__ T block_exclusive(T v, T *total, T *scratch /* 32 */)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T inc = v;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { T t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    T off = 0, tot = 0;
    for (int i = 0; i < SCAN_T / 32; ++i) { T c = scratch[i]; if (i < wid) off += c; tot += c; }
    __syncthreads();
    *total = tot;
    return off + inc - v;
}

template <class Tin, class Tout>
__global__ void __launch_bounds__(SCAN_T) scan_reduce_kernel(const Tin *in, uint64_t n, Tout *sums)
{
    __shared__ Tout scratch[32];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_B + threadIdx.x * SCAN_I;
    Tout s = 0;
    #pragma unroll
    for (int i = 0; i < SCAN_I; ++i) if (base + i < n) s += (Tout)in[base + i];
    Tout tot;
    block_exclusive<Tout>(s, &tot, scratch);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

template <class Tin, class Tout>
__global__ void __launch_bounds__(SCAN_T) scan_apply_kernel(const Tin *in, Tout *out, uint64_t n, const Tout *sums_scanned)
{
    __shared__ Tout scratch[32];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_B + threadIdx.x * SCAN_I;
    Tout v[SCAN_I]; Tout s = 0;
    #pragma unroll
    for (int i = 0; i < SCAN_I; ++i) { v[i] = base + i < n ? (Tout)in[base + i] : 0; s += v[i]; }
    Tout tot;
    Tout ex = block_exclusive<Tout>(s, &tot, scratch) + (sums_scanned ? sums_scanned[blockIdx.x] : 0);
    #pragma unroll
    for (int i = 0; i < SCAN_I; ++i) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
}

template <class Tout>
static size_t scan_scratch_elems(uint64_t n)
{
    size_t tot = 0;
    while (n > SCAN_B) { n = (n + SCAN_B - 1) / SCAN_B; tot += n; }
    return tot + 1;
}

template <class Tin, class Tout>
static cudaError_t scan_rec(const Tin *in, Tout *out, uint64_t n, Tout *scratch, cudaStream_t st, uint64_t *launches)
{
    if (n == 0) return cudaSuccess;
    uint64_t nb = (n + SCAN_B - 1) / SCAN_B;
    if (nb == 1) {
        scan_apply_kernel<Tin, Tout><<<1, SCAN_T, 0, st>>>(in, out, n, nullptr);
        PHI_LAUNCH_CHECK();
        return cudaSuccess;
    }
    Tout *sums = scratch;
    scan_reduce_kernel<Tin, Tout><<<(unsigned)nb, SCAN_T, 0, st>>>(in, n, sums);
    PHI_LAUNCH_CHECK();
    cudaError_t e = scan_rec<Tout, Tout>(sums, sums, nb, scratch + nb, st, launches);
    if (e != cudaSuccess) return e;
    scan_apply_kernel<Tin, Tout><<<(unsigned)nb, SCAN_T, 0, st>>>(in, out, n, sums);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

size_t scan_u32_to_u64_scratch(uint64_t n) { return scan_scratch_elems<uint64_t>(n) * 8; }
cudaError_t scan_u32_to_u64(const uint32_t *in, uint64_t *out, uint64_t n, void *scratch, cudaStream_t st, uint64_t *launches)
{
    return scan_rec<uint32_t, uint64_t>(in, out, n, (uint64_t *)scratch, st, launches);
}
size_t scan_u32_scratch(uint64_t n) { return scan_scratch_elems<uint32_t>(n) * 4; }
cudaError_t scan_u32_inplace(uint32_t *data, uint64_t n, void *scratch, cudaStream_t st, uint64_t *launches)
{
    return scan_rec<uint32_t, uint32_t>(data, data, n, (uint32_t *)scratch, st, launches);
}
cudaError_t scan_packed_steps(const PackedStep *in, uint64_t *out, uint64_t n, void *scratch, cudaStream_t st, uint64_t *launches)
{
    return scan_rec<PackedStep, uint64_t>(in, out, n, (uint64_t *)scratch, st, launches);
}
cudaError_t scan_u32(const uint32_t *in, uint32_t *out, uint64_t n, void *scratch, cudaStream_t st, uint64_t *launches)
{
    return scan_rec<uint32_t, uint32_t>(in, out, n, (uint32_t *)scratch, st, launches);
}

// ------------------------------------------------------------------ radix sort
constexpr int RS_T = 256, RS_I = 16, RS_B = RS_T * RS_I, RS_WARPS = RS_T / 32;
constexpr size_t RS_SMEM = (size_t)RS_B * 12;          // staged keys (8 B) + values (4 B)

__global__ void __launch_bounds__(RS_T) radix_hist_kernel(const uint64_t *keys, uint64_t n, int shift, uint32_t *hist, uint32_t nb)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint64_t base = (uint64_t)blockIdx.x * RS_B;
    #pragma unroll 4
    for (int i = 0; i < RS_I; ++i) {
        uint64_t idx = base + (uint64_t)i * RS_T + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

// One pass of the stable LSD sort.  Ranking: every warp owns a contiguous 512-key slice of the block's 4096 keys and ranks
// it round by round with __match_any_sync (equal digits -> one leader bumps the warp's counter); a 256-thread pass turns
// the per-warp counters into block-wide exclusive offsets.  Scatter: keys (and values) are first placed at their
// block-local sorted position in shared memory, then written out so that consecutive threads write consecutive
// addresses of one digit run (coalesced), instead of 32 scattered 8-byte stores per warp.
__global__ void __launch_bounds__(RS_T) radix_scatter_kernel(const uint64_t *keys_in, const uint32_t *vals_in, uint64_t *keys_out,
                                                           uint32_t *vals_out, uint64_t n, int shift, const uint32_t *hist, uint32_t nb)
{
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t *skey = (uint64_t *)rs_smem;                                  // [RS_B]
    uint32_t *sval = (uint32_t *)(rs_smem + (size_t)RS_B * 8);             // [RS_B] (only touched when vals_in != nullptr)
    __shared__ uint32_t wcount[RS_WARPS][257];
    __shared__ uint32_t gbase[256], dstart[257];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 257; i += RS_T) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const uint64_t blk0 = (uint64_t)blockIdx.x * RS_B;
    const uint64_t base = blk0 + (uint64_t)wid * (RS_I * 32);
    uint64_t key[RS_I]; uint16_t rk[RS_I];
    #pragma unroll
    for (int r = 0; r < RS_I; ++r) {
        uint64_t idx = base + r * 32 + lane;
        bool valid = idx < n;
        key[r] = valid ? keys_in[idx] : 0;
        uint32_t d = valid ? (uint32_t)((key[r] >> shift) & 255) : 256u;
        uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader) { old = wcount[wid][d]; wcount[wid][d] = old + __popc(peers); }
        old = __shfl_sync(0xFFFFFFFFu, old, leader);
        rk[r] = (uint16_t)(old + __popc(peers & lanemask_lt()));
        __syncwarp();
    }
    __syncthreads();
    {   // per digit: exclusive offsets of the warps inside the block, block total -> dstart (scanned below)
        uint32_t d = threadIdx.x, run = 0;
        #pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { uint32_t c = wcount[w][d]; wcount[w][d] = run; run += c; }
        gbase[d] = hist[(uint64_t)d * nb + blockIdx.x];
        dstart[d] = run;
    }
    __syncthreads();
    if (wid == 0) {                                                        // exclusive scan of the 256 digit totals (8 per lane)
        uint32_t v[8], sum = 0;
        #pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] = dstart[lane * 8 + i]; sum += v[i]; }
        uint32_t inc = sum;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
        uint32_t ex = inc - sum;
        #pragma unroll
        for (int i = 0; i < 8; ++i) { dstart[lane * 8 + i] = ex; ex += v[i]; }
        if (lane == 31) dstart[256] = inc;
    }
    __syncthreads();
    #pragma unroll
    for (int r = 0; r < RS_I; ++r) {
        uint64_t idx = base + r * 32 + lane;
        if (idx < n) {
            uint32_t d = (uint32_t)((key[r] >> shift) & 255);
            uint32_t loc = dstart[d] + wcount[wid][d] + rk[r];             // block-local sorted position
            skey[loc] = key[r];
            if (vals_in) sval[loc] = vals_in[idx];
        }
    }
    __syncthreads();
    const uint32_t cnt = dstart[256];
    for (uint32_t i = threadIdx.x; i < cnt; i += RS_T) {
        uint64_t k = skey[i];
        uint32_t d = (uint32_t)((k >> shift) & 255);
        uint32_t dst = gbase[d] + (i - dstart[d]);
        keys_out[dst] = k;
        if (vals_in) vals_out[dst] = sval[i];
    }
}

size_t radix_sort_scratch(uint64_t n)
{
    uint64_t nb = (n + RS_B - 1) / RS_B;
    return (size_t)(256 * nb) * 4 + scan_u32_scratch(256 * nb) + 64;
}

cudaError_t radix_sort_u64(uint64_t *keys_a, uint64_t *keys_b, uint32_t *vals_a, uint32_t *vals_b, uint64_t n, int bit_lo, int bit_hi,
                           void *scratch, cudaStream_t st, uint64_t *launches)
{
    if (n <= 1 || bit_hi <= bit_lo) return cudaSuccess;
    if (n >= (1ull << 32)) return cudaErrorInvalidValue;
    {
        // 48 KB staging + 9 KB static shared memory > the 48 KB default.  Set on every call: the attribute belongs to the device the
        // caller is on, and one process may drive several devices (one host thread each).
        {
            cudaError_t e = cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM);
            if (e != cudaSuccess) return e;
        }
    }
    const uint32_t nb = (uint32_t)((n + RS_B - 1) / RS_B);
    uint32_t *hist = (uint32_t *)scratch;
    void *scan_scr = (void *)(hist + (size_t)256 * nb);
    uint64_t *kin = keys_a, *kout = keys_b; uint32_t *vin = vals_a, *vout = vals_b;
    int passes = 0;
    for (int shift = bit_lo; shift < bit_hi; shift += 8, ++passes) {
        radix_hist_kernel<<<nb, RS_T, 0, st>>>(kin, n, shift, hist, nb);
        PHI_LAUNCH_CHECK();
        cudaError_t e = scan_u32_inplace(hist, (uint64_t)256 * nb, scan_scr, st, launches);
        if (e != cudaSuccess) return e;
        radix_scatter_kernel<<<nb, RS_T, RS_SMEM, st>>>(kin, vin, kout, vout, n, shift, hist, nb);
        PHI_LAUNCH_CHECK();
        uint64_t *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    if (passes & 1) {                                   // result currently in keys_b: bring it home
        cudaError_t e = cudaMemcpyAsync(keys_a, keys_b, n * 8, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
        if (vals_a) { e = cudaMemcpyAsync(vals_a, vals_b, n * 4, cudaMemcpyDeviceToDevice, st); if (e != cudaSuccess) return e; }
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------ radix sort, u32 keys with wide digits
// Same scheme as above for (u32 key, u32 value) records with digits of up to 11 bits: a 20-bit key (the hash ranks of the
// anchors) is sorted in 2 passes instead of 3, and every pass moves 8 instead of 12 bytes per record.
constexpr int R32_MAXBINS = 2048;
constexpr size_t R32_SMEM = (size_t)RS_WARPS * (R32_MAXBINS + 2) * 2 + (size_t)RS_B * 8 + (size_t)(2 * R32_MAXBINS + 2) * 4;

__global__ void __launch_bounds__(RS_T) radix32_hist_kernel(const uint32_t *keys, uint64_t n, int shift, uint32_t nbins, uint32_t *hist, uint32_t nb)
{
    __shared__ uint32_t h[R32_MAXBINS];
    for (uint32_t i = threadIdx.x; i < nbins; i += RS_T) h[i] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_B;
    #pragma unroll 4
    for (int i = 0; i < RS_I; ++i) {
        uint64_t idx = base + (uint64_t)i * RS_T + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & (nbins - 1)], 1u);
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < nbins; d += RS_T) hist[(uint64_t)d * nb + blockIdx.x] = h[d];
}

__global__ void __launch_bounds__(RS_T) radix32_scatter_kernel(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                                                             uint64_t n, int shift, uint32_t nbins, const uint32_t *hist, uint32_t nb)
{
    extern __shared__ __align__(16) unsigned char r32_smem[];
    uint16_t *wcount = (uint16_t *)r32_smem;                               // [RS_WARPS][nbins + 1] keys of a digit seen by a warp (<= 512)
    const uint32_t wstride = nbins + 2;
    uint32_t *skey = (uint32_t *)(r32_smem + (size_t)RS_WARPS * (R32_MAXBINS + 2) * 2);
    uint32_t *sval = skey + RS_B;
    uint32_t *gbase = sval + RS_B, *dstart = gbase + R32_MAXBINS;          // [nbins], [nbins + 1]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < RS_WARPS * wstride; i += RS_T) wcount[i] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_B + (uint64_t)wid * (RS_I * 32);
    uint32_t key[RS_I]; uint16_t rk[RS_I];
    uint16_t *wc = wcount + (size_t)wid * wstride;
    #pragma unroll
    for (int r = 0; r < RS_I; ++r) {
        uint64_t idx = base + r * 32 + lane;
        bool valid = idx < n;
        key[r] = valid ? keys_in[idx] : 0;
        uint32_t d = valid ? ((key[r] >> shift) & (nbins - 1)) : nbins;
        uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader) { old = wc[d]; wc[d] = (uint16_t)(old + __popc(peers)); }
        old = __shfl_sync(0xFFFFFFFFu, old, leader);
        rk[r] = (uint16_t)(old + __popc(peers & lanemask_lt()));
        __syncwarp();
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < nbins; d += RS_T) {                 // exclusive offsets of the warps inside the block, per digit
        uint32_t run = 0;
        #pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { uint32_t c = wcount[(size_t)w * wstride + d]; wcount[(size_t)w * wstride + d] = (uint16_t)run; run += c; }
        gbase[d] = hist[(uint64_t)d * nb + blockIdx.x];
        dstart[d] = run;
    }
    __syncthreads();
    if (wid == 0) {                                                        // exclusive scan of the digit totals (nbins / 32 per lane)
        const uint32_t per = nbins >> 5 ? nbins >> 5 : 1;
        uint32_t sum = 0;
        for (uint32_t i = 0; i < per; ++i) { uint32_t j = lane * per + i; if (j < nbins) sum += dstart[j]; }
        uint32_t inc = sum;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
        uint32_t ex = inc - sum;
        for (uint32_t i = 0; i < per; ++i) { uint32_t j = lane * per + i; if (j < nbins) { uint32_t v = dstart[j]; dstart[j] = ex; ex += v; } }
        if (lane == 31) dstart[nbins] = inc;
    }
    __syncthreads();
    #pragma unroll
    for (int r = 0; r < RS_I; ++r) {
        uint64_t idx = base + r * 32 + lane;
        if (idx < n) {
            uint32_t d = (key[r] >> shift) & (nbins - 1);
            uint32_t loc = dstart[d] + wc[d] + rk[r];                      // block-local sorted position
            skey[loc] = key[r];
            sval[loc] = vals_in[idx];
        }
    }
    __syncthreads();
    const uint32_t cnt = dstart[nbins];
    for (uint32_t i = threadIdx.x; i < cnt; i += RS_T) {
        uint32_t k = skey[i];
        uint32_t d = (k >> shift) & (nbins - 1);
        uint32_t dst = gbase[d] + (i - dstart[d]);
        keys_out[dst] = k;
        vals_out[dst] = sval[i];
    }
}

size_t radix_sort_u32_scratch(uint64_t n)
{
    uint64_t nb = (n + RS_B - 1) / RS_B;
    return (size_t)(R32_MAXBINS * nb) * 4 + scan_u32_scratch(R32_MAXBINS * nb) + 64;
}

cudaError_t radix_sort_u32(uint32_t *keys_a, uint32_t *keys_b, uint32_t *vals_a, uint32_t *vals_b, uint64_t n, int bits,
                           void *scratch, cudaStream_t st, uint64_t *launches)
{
    if (n <= 1 || bits <= 0) return cudaSuccess;
    if (n >= (1ull << 32) || bits > 32) return cudaErrorInvalidValue;
    {
        {
            cudaError_t e = cudaFuncSetAttribute(radix32_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R32_SMEM);
            if (e != cudaSuccess) return e;
        }
    }
    const int passes = (bits + 10) / 11;                                   // digits of at most 11 bits, as even as possible
    const int dbits = (bits + passes - 1) / passes;
    const uint32_t nbins = 1u << dbits;
    const uint32_t nb = (uint32_t)((n + RS_B - 1) / RS_B);
    uint32_t *hist = (uint32_t *)scratch;
    void *scan_scr = (void *)(hist + (size_t)nbins * nb);
    uint32_t *kin = keys_a, *kout = keys_b, *vin = vals_a, *vout = vals_b;
    for (int p = 0; p < passes; ++p) {
        const int shift = p * dbits;
        radix32_hist_kernel<<<nb, RS_T, 0, st>>>(kin, n, shift, nbins, hist, nb);
        PHI_LAUNCH_CHECK();
        cudaError_t e = scan_u32_inplace(hist, (uint64_t)nbins * nb, scan_scr, st, launches);
        if (e != cudaSuccess) return e;
        radix32_scatter_kernel<<<nb, RS_T, R32_SMEM, st>>>(kin, vin, kout, vout, n, shift, nbins, hist, nb);
        PHI_LAUNCH_CHECK();
        uint32_t *t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    if (passes & 1) {                                                      // result currently in keys_b / vals_b: bring it home
        cudaError_t e = cudaMemcpyAsync(keys_a, keys_b, n * 4, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(vals_a, vals_b, n * 4, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------ ordered spectrum table -> sorted array
// A probe cluster is a maximal run of occupied slots.  Every key of a cluster has its home slot inside the cluster, homes are
// monotone in the key, and clusters are separated by an EMPTY slot: sorting each cluster in place sorts the table.
// One thread per cluster head; clusters are short (load <= ~50 %, uniform hashes), so this is a few compares per key.
// The same pass counts the occupied slots of every TABLE_BLOCK slots (sorting moves keys inside a cluster only: which slots
// are occupied does not change), so the table is streamed once for both.
__global__ void __launch_bounds__(256) table_cluster_sort_kernel(uint64_t *table, uint64_t limit, uint32_t *block_cnt)
{
    __shared__ uint32_t scratch[32];
    constexpr int I = (int)(TABLE_BLOCK / 256);
    const uint64_t base = (uint64_t)blockIdx.x * TABLE_BLOCK + threadIdx.x;       // slot j = base + i * 256: coalesced
    uint32_t c = 0, heads = 0;
    #pragma unroll
    for (int i = 0; i < I; ++i) {
        const uint64_t j = base + (uint64_t)i * 256;
        const uint64_t v = j < limit ? table[j] : TABLE_EMPTY;
        uint64_t left = __shfl_up_sync(0xFFFFFFFFu, v, 1);
        if ((threadIdx.x & 31) == 0) left = (j && j <= limit) ? table[j - 1] : TABLE_EMPTY;
        const bool occ = v != TABLE_EMPTY;
        c += occ;
        if (occ && left == TABLE_EMPTY) heads |= 1u << i;                         // cluster head
    }
    uint32_t tot;
    block_exclusive<uint32_t>(c, &tot, scratch);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = tot;
    #pragma unroll
    for (int i = 0; i < I; ++i) {
        if (!((heads >> i) & 1u)) continue;
        const uint64_t s = base + (uint64_t)i * 256;
        for (uint64_t a = s + 1; table[a] != TABLE_EMPTY; ++a) {                  // table[limit] is EMPTY: stops there at the latest
            uint64_t x = table[a], b = a;
            while (b > s && table[b - 1] > x) { table[b] = table[b - 1]; --b; }
            table[b] = x;
        }
    }
}

__global__ void __launch_bounds__(256) table_write_kernel(const uint64_t *table, uint64_t limit, const uint32_t *block_off, uint64_t *out)
{
    __shared__ uint32_t scratch[32];
    constexpr int I = (int)(TABLE_BLOCK / 256);
    uint64_t base = (uint64_t)blockIdx.x * TABLE_BLOCK + (uint64_t)threadIdx.x * I;   // consecutive slots per thread: slot order is kept
    uint64_t v[I]; uint32_t c = 0;
    #pragma unroll
    for (int i = 0; i < I; ++i) { v[i] = base + i < limit ? table[base + i] : TABLE_EMPTY; c += v[i] != TABLE_EMPTY; }
    uint32_t tot;
    uint32_t ex = block_exclusive<uint32_t>(c, &tot, scratch);
    uint64_t o = (uint64_t)block_off[blockIdx.x] + ex;
    #pragma unroll
    for (int i = 0; i < I; ++i) if (v[i] != TABLE_EMPTY) out[o++] = v[i];
}

// insert a key array (owner side of the multi-GPU spectrum exchange): home slot = umulhi(key - base, mult), upwards probing
__global__ void table_insert_keys_kernel(const uint64_t *keys, uint64_t n, uint64_t *table, uint64_t base, uint64_t mult, uint64_t limit,
                                         unsigned long long *ctr)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = keys[i];
    if (key == TABLE_EMPTY) { ctr[CTR_HAS_MAXKEY] = 1; return; }
    for (uint64_t slot = __umul64hi(key - base, mult); slot < limit; ++slot) {
        uint64_t cur = table[slot];
        if (cur == key) return;
        if (cur == TABLE_EMPTY) {
            uint64_t old = atomicCAS((unsigned long long *)&table[slot], (unsigned long long)TABLE_EMPTY, (unsigned long long)key);
            if (old == TABLE_EMPTY || old == key) return;
        }
    }
    ctr[CTR_OVERFLOW] = 1;
}

cudaError_t table_insert_keys(const uint64_t *keys, uint64_t n, uint64_t *table, uint64_t base, uint64_t mult, uint64_t limit,
                              unsigned long long *ctr, cudaStream_t st, uint64_t *launches)
{
    if (!n) return cudaSuccess;
    table_insert_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys, n, table, base, mult, limit, ctr);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

size_t table_blocks(uint64_t limit) { return (size_t)((limit + TABLE_BLOCK - 1) / TABLE_BLOCK); }

cudaError_t table_sort_and_count(uint64_t *table, uint64_t limit, uint32_t *block_cnt, cudaStream_t st, uint64_t *launches)
{
    if (!limit) return cudaSuccess;
    table_cluster_sort_kernel<<<(unsigned)table_blocks(limit), 256, 0, st>>>(table, limit, block_cnt);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t table_write_ordered(const uint64_t *table, uint64_t limit, const uint32_t *block_off, uint64_t *out, cudaStream_t st, uint64_t *launches)
{
    if (!limit) return cudaSuccess;
    table_write_kernel<<<(unsigned)table_blocks(limit), 256, 0, st>>>(table, limit, block_off, out);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

// ------------------------------------------------------------------ radix directory
__global__ void directory_kernel(const uint64_t *sorted, uint32_t n, int dbits, uint32_t *dir)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t nbuckets = 1ull << dbits;
    uint64_t pb = dbits ? sorted[i] >> (64 - dbits) : 0;
    long long pprev = i ? (long long)(dbits ? sorted[i - 1] >> (64 - dbits) : 0) : -1;
    for (long long b = pprev + 1; b <= (long long)pb; ++b) dir[b] = i;
    if (i == n - 1) for (uint64_t b = pb + 1; b <= nbuckets; ++b) dir[b] = n;
}

cudaError_t build_directory(const uint64_t *sorted, uint32_t n, int dbits, uint32_t *dir, cudaStream_t st, uint64_t *launches)
{
    if (n == 0) return fill_u32(dir, (1ull << dbits) + 1, 0, st, launches);
    directory_kernel<<<(n + 255) / 256, 256, 0, st>>>(sorted, n, dbits, dir);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

// ------------------------------------------------------------------ fills
template <class T>
__global__ void fill_kernel(T *p, uint64_t n, T v)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
cudaError_t fill_u64(uint64_t *p, uint64_t n, uint64_t v, cudaStream_t st, uint64_t *launches)
{
    if (!n) return cudaSuccess;
    unsigned nb = (unsigned)((n + 1023) / 1024 < 148 * 16 ? (n + 1023) / 1024 : 148 * 16);
    fill_kernel<uint64_t><<<nb, 256, 0, st>>>(p, n, v);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}
cudaError_t fill_u32(uint32_t *p, uint64_t n, uint32_t v, cudaStream_t st, uint64_t *launches)
{
    if (!n) return cudaSuccess;
    unsigned nb = (unsigned)((n + 1023) / 1024 < 148 * 16 ? (n + 1023) / 1024 : 148 * 16);
    fill_kernel<uint32_t><<<nb, 256, 0, st>>>(p, n, v);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

}  // namespace phi
