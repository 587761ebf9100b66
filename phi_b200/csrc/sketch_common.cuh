// Helpers shared by the read-sketch and walk-sketch translation units (each is compiled for its own tile size).
#pragma once
#include "sketch_tile.cuh"

namespace phi {
namespace {

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

}  // namespace
}  // namespace phi
