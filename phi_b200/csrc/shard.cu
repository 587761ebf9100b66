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

// ---------------------------------------------------------------------------------------------------------------------
// Region partition of the walks.  Splitting the walks BY WALK over the GPUs would undo walk sharing (chunks.cu): identical
// chunks of different walks would land on different GPUs and be sketched once per GPU.  So every GPU gets ALL walks, cut to the
// steps whose vertices lie in its range of the topological base coordinate (bases of all vertices that precede a vertex in
// top_order_map), plus the context the owned windows need: >= w bases in front (the k-mers of the first owned window and the
// window before it, /root/reference/src/ILP_index.cpp:405-414) and >= k-1 bases behind (the last owned k-mer).  The library then
// owns exactly the windows whose last k-mer starts on a vertex inside the range (phi_gpu_index_set_walk_region), every window
// of every walk is owned by exactly one GPU, and the union of the per-GPU results is the reference's result
// (phi_index_result_merge).
#include <algorithm>
#include <vector>

namespace {

// coordinate of every vertex, or false when top_order_map is not a permutation of [0, n_vtx)
bool topo_coordinates(const phi_graph_view *g, std::vector<uint64_t> &coord)
{
    const uint32_t V = g->n_vtx;
    std::vector<uint64_t> len_at(V, 0);
    std::vector<uint8_t> seen(V, 0);
    for (uint32_t v = 0; v < V; ++v) {
        const int32_t t = g->top_order_map[v];
        if (t < 0 || (uint32_t)t >= V || seen[t]) return false;
        seen[t] = 1;
        len_at[t] = g->seg_off[v + 1] - g->seg_off[v];
    }
    uint64_t run = 0;
    for (uint32_t t = 0; t < V; ++t) { const uint64_t l = len_at[t]; len_at[t] = run; run += l; }
    coord.resize(V);
    for (uint32_t v = 0; v < V; ++v) coord[v] = len_at[g->top_order_map[v]];
    return true;
}

}  // namespace

extern "C" int phi_shard_walk_regions(const phi_graph_view *g, int world, uint64_t *coord_bounds)
{
    if (!g || !coord_bounds || world < 1) return PHI_ERR_ARG;
    coord_bounds[0] = 0; coord_bounds[world] = ~0ull;
    for (int r = 1; r < world; ++r) coord_bounds[r] = ~0ull;
    if (world == 1 || !g->n_vtx) return PHI_OK;
    std::vector<uint64_t> coord;
    if (!topo_coordinates(g, coord)) return PHI_ERR_UNSUPPORTED;
    const uint64_t total = g->seg_off[g->n_vtx];
    // steps per coordinate bin over the given walks (a sample of the walks is enough), equal-weight cuts
    const int BINS = 1 << 14;
    std::vector<uint64_t> hist(BINS + 1, 0);
    const uint64_t S = g->n_walks ? g->walk_off[g->n_walks] : 0;
    const unsigned __int128 scale = total ? total : 1;
    for (uint64_t s = 0; s < S; ++s) hist[(size_t)(((unsigned __int128)coord[g->walk_vtx[s]] * BINS) / scale)]++;
    if (!S) for (int b = 0; b < BINS; ++b) hist[b] = 1;
    uint64_t sum = 0; for (int b = 0; b < BINS; ++b) sum += hist[b];
    uint64_t run = 0; int b = 0;
    for (int r = 1; r < world; ++r) {
        const uint64_t target = (uint64_t)(((unsigned __int128)sum * (unsigned)r) / (unsigned)world);
        while (b < BINS && run + hist[b] <= target) run += hist[b++];
        coord_bounds[r] = (uint64_t)(((unsigned __int128)total * (unsigned)b) / BINS);
    }
    for (int r = 1; r <= world; ++r) if (coord_bounds[r] < coord_bounds[r - 1]) coord_bounds[r] = coord_bounds[r - 1];
    return PHI_OK;
}

extern "C" int phi_shard_slice_walks(const phi_graph_view *g, int k, int w, uint64_t coord_lo, uint64_t coord_hi,
                                     uint64_t *slice_first, uint64_t *slice_len)
{
    if (!g || !slice_first || !slice_len || k < 1 || w < 1) return PHI_ERR_ARG;
    std::vector<uint64_t> coord;
    if (!topo_coordinates(g, coord)) return PHI_ERR_UNSUPPORTED;
    for (uint32_t h = 0; h < g->n_walks; ++h) {
        const uint64_t s0 = g->walk_off[h], s1 = g->walk_off[h + 1];
        const uint32_t *wv = g->walk_vtx;
        for (uint64_t s = s0; s + 1 < s1; ++s)
            if (coord[wv[s]] > coord[wv[s + 1]]) return PHI_ERR_UNSUPPORTED;     // a walk that does not follow the topological order: no region cut
        // a = first step with coordinate >= lo, b = first step with coordinate >= hi
        uint64_t a = s0, b = s1, x = s0, y = s1;
        while (x < y) { const uint64_t m = (x + y) >> 1; if (coord[wv[m]] < coord_lo) x = m + 1; else y = m; }
        a = x; y = s1;
        while (x < y) { const uint64_t m = (x + y) >> 1; if (coord[wv[m]] < coord_hi) x = m + 1; else y = m; }
        b = x;
        if (a >= b) { slice_first[h] = a; slice_len[h] = 0; continue; }
        uint64_t L = a, have = 0;
        while (L > s0 && have < (uint64_t)w) { --L; have += g->seg_off[wv[L] + 1] - g->seg_off[wv[L]]; }
        uint64_t R = b; have = 0;
        while (R < s1 && have < (uint64_t)(k - 1)) { have += g->seg_off[wv[R] + 1] - g->seg_off[wv[R]]; ++R; }
        slice_first[h] = L; slice_len[h] = R - L;
    }
    return PHI_OK;
}
