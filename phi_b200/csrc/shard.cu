// Host-only partition helpers of the C ABI (no GPU needed) — see include/phi_gpu_index.h.
// The reference is single-process OpenMP (/root/reference/src/ILP_index.cpp:545,559,617,674); the multi-GPU
// partition is new: reads and walks are split into contiguous, weight-balanced shards, and read-minimizer
// hashes are owned by GPU floor(hash * world / 2^64) (range partition on the high bits, so that the
// concatenation of the per-owner sorted spectra is the globally sorted spectrum and
// global rank = local rank + exclusive prefix of the per-owner distinct counts).
#include "../../include/phi_gpu_index.h"

extern "C" int phi_shard_owner_of_hash(uint64_t hash, int world)
{
    if (world <= 1) return 0;
    return (int)(((unsigned __int128)hash * (unsigned __int128)(uint64_t)world) >> 64);
}

extern "C" int phi_shard_split_by_weight(const uint64_t *off, uint64_t n, int world, uint64_t *bounds)
{
    if (!bounds || world < 1 || (n && !off)) return PHI_ERR_ARG;
    const uint64_t base = n ? off[0] : 0, total = n ? off[n] - base : 0;
    bounds[0] = 0;
    uint64_t i = 0;
    for (int r = 1; r < world; ++r) {
        // first item whose start offset reaches r/world of the total weight
        unsigned __int128 target = (unsigned __int128)total * (unsigned)r / (unsigned)world;
        while (i < n && (unsigned __int128)(off[i] - base) < target) ++i;
        bounds[r] = i;
    }
    bounds[world] = n;
    return PHI_OK;
}
