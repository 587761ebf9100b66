// The host side of a multi-GPU front end run, independent of the reference's classes: the flat views of the whole input cut into one
// shard per GPU.  Reads: contiguous, balanced by bases (they point into the caller's arrays).  Walks: by REGION of the topological
// base coordinate (every GPU gets all walks cut to its range plus context, so that walk sharing keeps working across GPUs:
// phi_shard_walk_regions -> phi_shard_slice_walks_all -> one copy of the slices per GPU); by whole walks (pointing into the caller's
// array) when the graph does not allow a region cut.  The reference has no counterpart: it is single-process OpenMP
// (/root/reference/src/ILP_index.cpp:545,559,617,674).  Used by phi_adapter.hpp; tests/test_adapter_shards.py drives it on the CPU.
#ifndef PHI_SHARDS_HPP
#define PHI_SHARDS_HPP

#include "phi_gpu_index.h"

#include <cstdint>
#include <cstring>
#include <vector>

namespace phi_adapter {
namespace detail {

struct Shard {                                     // the views of one GPU: its walks (whole walks or region slices) and reads, offsets rebased to 0
    std::vector<uint64_t> walk_off, read_off;
    std::vector<uint32_t> walk_vtx;                // region slices only (by-walk shards point into the caller's array)
    phi_graph_view g; phi_reads_view rd;
    uint32_t walk_id_base;
    bool region; uint64_t coord_lo, coord_hi;
};

// shards[r] = the input of GPU r of W.  `g` and `rd` must stay valid while the shards are used.  Returns true when the walks were cut
// by region (every shard: all walks, walk_id_base 0, its coordinate range), false for whole walks (walk_id_base = first walk).
inline bool build_shards(const phi_graph_view &g, const phi_reads_view &rd, int W, int k, int w, std::vector<Shard> &shards)
{
    shards.assign((size_t)W, Shard());
    const size_t NWK = (size_t)g.n_walks;
    std::vector<uint64_t> wb((size_t)W + 1, 0), rb((size_t)W + 1, 0), cb((size_t)W + 1, 0);
    wb[W] = g.n_walks; rb[W] = rd.n_reads;
    static const uint64_t zero_off[1] = {0};
    const uint64_t *walk_off = g.n_walks ? g.walk_off : zero_off, *read_off = rd.n_reads ? rd.read_off : zero_off;
    bool by_region = false;
    if (W > 1) {
        phi_shard_split_by_weight(read_off, rd.n_reads, W, rb.data());
        by_region = phi_shard_walk_regions(&g, W, cb.data()) == PHI_OK;
    }
    std::vector<uint64_t> sl_first((size_t)W * NWK), sl_len((size_t)W * NWK);
    if (by_region) by_region = phi_shard_slice_walks_all(&g, k, w, W, cb.data(), sl_first.data(), sl_len.data()) == PHI_OK;
    if (by_region) {
        // every GPU's copy of its slices: offsets first, then the (GPU, walk) copies in parallel
        for (int r = 0; r < W; ++r) {
            Shard &s = shards[r];
            s.walk_off.assign(NWK + 1, 0);
            for (size_t h = 0; h < NWK; ++h) s.walk_off[h + 1] = s.walk_off[h] + sl_len[(size_t)r * NWK + h];
            s.walk_vtx.resize(s.walk_off[NWK]);
            s.region = true; s.coord_lo = cb[r]; s.coord_hi = cb[r + 1]; s.walk_id_base = 0;
        }
#pragma omp parallel for schedule(dynamic, 1)
        for (long long t = 0; t < (long long)((size_t)W * NWK); ++t) {
            const size_t r = (size_t)t / NWK, h = (size_t)t % NWK;
            if (sl_len[t]) memcpy(shards[r].walk_vtx.data() + shards[r].walk_off[h], g.walk_vtx + sl_first[t], (size_t)sl_len[t] * sizeof(uint32_t));
        }
        for (int r = 0; r < W; ++r) { Shard &s = shards[r]; s.g = g; s.g.walk_off = s.walk_off.data(); s.g.walk_vtx = s.walk_vtx.data(); }
    } else {
        if (W > 1) phi_shard_split_by_weight(walk_off, g.n_walks, W, wb.data());
        for (int r = 0; r < W; ++r) {
            Shard &s = shards[r];
            for (uint64_t h = wb[r]; h <= wb[r + 1]; ++h) s.walk_off.push_back(walk_off[h] - walk_off[wb[r]]);
            s.region = false; s.coord_lo = 0; s.coord_hi = ~0ull; s.walk_id_base = (uint32_t)wb[r];
            s.g = g; s.g.n_walks = (uint32_t)(wb[r + 1] - wb[r]); s.g.walk_off = s.walk_off.data();
            s.g.walk_vtx = g.walk_vtx ? g.walk_vtx + walk_off[wb[r]] : g.walk_vtx;
        }
    }
    for (int r = 0; r < W; ++r) {
        Shard &s = shards[r];
        for (uint64_t q = rb[r]; q <= rb[r + 1]; ++q) s.read_off.push_back(read_off[q] - read_off[rb[r]]);
        s.rd.n_reads = rb[r + 1] - rb[r]; s.rd.read_off = s.read_off.data();
        s.rd.read_bases = rd.read_bases ? rd.read_bases + read_off[rb[r]] : rd.read_bases;
    }
    return by_region;
}

}  // namespace detail
}  // namespace phi_adapter
#endif
