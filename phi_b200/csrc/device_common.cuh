// Device-side building blocks shared by the sketch kernels (sm_100a).
//
//   murmur3_x64_128_xor  — MurmurHash3_x64_128(seed 0) -> h[0]^h[1], the reference's
//                          hash128_to_64 (/root/reference/src/ILP_index.cpp:10-18;
//                          /root/reference/src/MurmurHash3.cpp:255-332).
//   2-bit k-mer helpers  — A<C<G<T == 0<1<2<3 so that unsigned integer order on the
//                          packed value equals std::string order on the upper-cased
//                          k-mer (/root/reference/src/ILP_index.cpp:394, :397).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace phi {

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

__device__ __forceinline__ uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

// Words W[0..3] hold the key bytes little-endian (byte i of the key in bits 8*(i&7) of W[i>>3]);
// bytes at index >= len MUST be zero.  len in [1, 32].  len is warp-uniform in every caller.
__device__ __forceinline__ uint64_t murmur3_x64_128_xor(const uint64_t W[4], int len)
{
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = 0, h2 = 0;
    const int nblocks = len >> 4;
    #pragma unroll
    for (int b = 0; b < 2; ++b) {
        if (b < nblocks) {
            uint64_t k1 = W[2 * b], k2 = W[2 * b + 1];
            k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
            h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
            k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
            h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
        }
    }
    const int rem = len & 15;
    if (rem) {
        uint64_t t1 = nblocks == 0 ? W[0] : W[2], t2 = nblocks == 0 ? W[1] : W[3];
        if (rem > 8) { t2 *= c2; t2 = rotl64(t2, 33); t2 *= c1; h2 ^= t2; }
        t1 *= c1; t1 = rotl64(t1, 31); t1 *= c2; h1 ^= t1;
    }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    return h1 ^ h2;
}

// The same hash over a key of any length whose bytes come from a functor (byte(i), i in [0, len)): k-mers longer than 32 bases and
// k-mers with non-ACGT bytes are hashed from their spelling (/root/reference/src/MurmurHash3.cpp:255-332: 16-byte blocks as two
// little-endian u64, then the tail switch, then the finalisation).
template <class ByteAt>
__device__ __forceinline__ uint64_t murmur3_x64_128_xor_bytes(ByteAt byte, int len)
{
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = 0, h2 = 0;
    const int nblocks = len >> 4;
    for (int b = 0; b < nblocks; ++b) {
        uint64_t k1 = 0, k2 = 0;
        for (int j = 0; j < 8; ++j) { k1 |= (uint64_t)byte(16 * b + j) << (8 * j); k2 |= (uint64_t)byte(16 * b + 8 + j) << (8 * j); }
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const int rem = len & 15;
    if (rem) {
        uint64_t t1 = 0, t2 = 0;
        for (int j = 0; j < rem && j < 8; ++j) t1 |= (uint64_t)byte(16 * nblocks + j) << (8 * j);
        for (int j = 8; j < rem; ++j) t2 |= (uint64_t)byte(16 * nblocks + j) << (8 * (j - 8));
        if (rem > 8) { t2 *= c2; t2 = rotl64(t2, 33); t2 *= c1; h2 ^= t2; }
        t1 *= c1; t1 = rotl64(t1, 31); t1 *= c2; h1 ^= t1;
    }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    return h1 ^ h2;
}

// ::toupper in the C locale (/root/reference/src/ILP_index.cpp:369, :449)
__device__ __forceinline__ uint32_t upcase(uint32_t c) { return (c - 'a' < 26u) ? c - 32 : c; }
// upper-cased byte -> is it one of A C G T
__device__ __forceinline__ bool is_acgt(uint32_t c) { return ((c & 0xE0u) == 0x40u) && ((0x0010008Au >> (c & 31u)) & 1u); }
// A,C,G,T (either case) -> 0,1,2,3
__device__ __forceinline__ uint32_t code2(uint32_t c) { return ((c >> 1) ^ (c >> 2)) & 3u; }
// reverse_strand on one upper-cased byte (/root/reference/src/ILP_index.cpp:335-353)
__device__ __forceinline__ uint32_t comp_byte(uint32_t c)
{
    return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c;
}

// reverse the order of the 32 two-bit groups of x
__device__ __forceinline__ uint64_t rev2(uint64_t x)
{
    uint32_t lo = __brev((uint32_t)(x >> 32)), hi = __brev((uint32_t)x);
    lo = ((lo >> 1) & 0x55555555u) | ((lo & 0x55555555u) << 1);
    hi = ((hi >> 1) & 0x55555555u) | ((hi & 0x55555555u) << 1);
    return ((uint64_t)hi << 32) | lo;
}

// reverse complement of a k-mer packed right-aligned (first base in the top 2 bits of the 2k-bit field)
__device__ __forceinline__ uint64_t revcomp2(uint64_t fwd, int k) { return rev2(~fwd) >> (64 - 2 * k); }

// 8 two-bit codes (code i in bits 2i..2i+1 of x16) -> 8 ASCII bytes (byte i = base i)
__device__ __forceinline__ uint64_t codes8_to_ascii(uint32_t x16)
{
    uint32_t x = x16 & 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;                // nibble i = code i
    const uint32_t tbl = 0x54474341u;                // 'A','C','G','T'
    uint32_t lo = __byte_perm(tbl, 0, x & 0xFFFFu), hi = __byte_perm(tbl, 0, x >> 16);
    return ((uint64_t)hi << 32) | lo;
}

// hash128_to_64 of the ASCII spelling of a packed k-mer (right-aligned, first base on top)
__device__ __forceinline__ uint64_t hash_packed_kmer(uint64_t km, int k)
{
    uint64_t r = rev2(km << (64 - 2 * k));           // base i now in bits 2i..2i+1
    uint64_t W[4];
    #pragma unroll
    for (int j = 0; j < 4; ++j) {
        int nb = k - 8 * j;                          // bytes of the key living in word j
        uint64_t wv = codes8_to_ascii((uint32_t)(r >> (16 * j)));
        W[j] = nb >= 8 ? wv : nb <= 0 ? 0ull : (wv & ((1ull << (8 * nb)) - 1));
    }
    return murmur3_x64_128_xor(W, k);
}

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace phi
