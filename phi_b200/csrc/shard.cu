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
        // first item whose start offset reaches r/world of the total weight (offsets ascend: binary search from the last cut on)
        const unsigned __int128 target = (unsigned __int128)total * (unsigned)r / (unsigned)world;
        uint64_t lo = i, hi = n;
        while (lo < hi) { const uint64_t m = lo + ((hi - lo) >> 1); if ((unsigned __int128)(off[m] - base) < target) lo = m + 1; else hi = m; }
        bounds[r] = i = lo;
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
#include <atomic>
#include <cstdlib>
#include <thread>
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

// fn(task) for task in [0, n) on up to 16 host threads (PHI_SHARD_THREADS overrides); `work` = items behind the tasks, small jobs stay
// on the calling thread.
template <class F> void shard_parallel(size_t n, uint64_t work, F fn)
{
    int threads = 1;
    if (const char *e = getenv("PHI_SHARD_THREADS")) threads = std::max(1, std::min(64, atoi(e)));
    else if (work >= (1u << 20)) threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    threads = (int)std::min<size_t>((size_t)threads, n);
    if (threads <= 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
    std::atomic<size_t> next(0);
    auto body = [&]() { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(body);
    body();
    for (std::thread &t : pool) t.join();
}

// the walk steps cut into pieces of at most PIECE steps (a piece never spans two walks): the unit of the parallel passes
struct StepPiece { uint32_t walk; uint64_t s0, s1; };
std::vector<StepPiece> step_pieces(const phi_graph_view *g)
{
    uint64_t PIECE = 1u << 21;
    if (const char *e = getenv("PHI_SHARD_PIECE")) PIECE = std::max<uint64_t>(1, strtoull(e, nullptr, 10));   // tests: piece borders everywhere
    std::vector<StepPiece> out;
    for (uint32_t h = 0; h < g->n_walks; ++h)
        for (uint64_t s = g->walk_off[h]; s < g->walk_off[h + 1]; s += PIECE) out.push_back({h, s, std::min<uint64_t>(s + PIECE, g->walk_off[h + 1])});
    return out;
}

// every walk follows the topological order (non-decreasing coordinate from step to step)?
bool walks_follow_order(const phi_graph_view *g, const std::vector<uint64_t> &coord)
{
    const std::vector<StepPiece> pieces = step_pieces(g);
    const uint32_t *wv = g->walk_vtx;
    std::atomic<int> bad(0);
    shard_parallel(pieces.size(), g->n_walks ? g->walk_off[g->n_walks] : 0, [&](size_t i) {
        if (bad.load(std::memory_order_relaxed)) return;
        const StepPiece &pc = pieces[i];
        const uint64_t end = std::min<uint64_t>(pc.s1 + 1, g->walk_off[pc.walk + 1]);     // the pair across the piece border belongs to this piece
        if (pc.s0 >= end) return;
        uint64_t prev = coord[wv[pc.s0]];
        for (uint64_t s = pc.s0 + 1; s < end; ++s) { const uint64_t c = coord[wv[s]]; if (c < prev) { bad.store(1); return; } prev = c; }
    });
    return !bad.load();
}

// slices of all walks for the region [lo, hi): see phi_shard_slice_walks
void slice_region(const phi_graph_view *g, const std::vector<uint64_t> &coord, int k, int w, uint64_t coord_lo, uint64_t coord_hi,
                  uint64_t *slice_first, uint64_t *slice_len)
{
    const uint32_t *wv = g->walk_vtx;
    for (uint32_t h = 0; h < g->n_walks; ++h) {
        const uint64_t s0 = g->walk_off[h], s1 = g->walk_off[h + 1];
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
    // steps per coordinate bin over the given walks, equal-weight cuts.  The bin of a vertex is computed once per vertex; big walk
    // sets are sampled (every stride-th step, at most ~8 M samples: the cuts move by a fraction of a bin) piece by piece on
    // several threads.
    const int BINS = 1 << 14;
    const unsigned __int128 scale = total ? total : 1;
    std::vector<uint16_t> bin_of(g->n_vtx);
    for (uint32_t v = 0; v < g->n_vtx; ++v) bin_of[v] = (uint16_t)std::min<uint64_t>(BINS - 1, (uint64_t)(((unsigned __int128)coord[v] * BINS) / scale));
    const uint64_t S = g->n_walks ? g->walk_off[g->n_walks] : 0;
    const uint64_t stride = std::max<uint64_t>(1, S >> 23);
    const std::vector<StepPiece> pieces = step_pieces(g);
    std::vector<std::vector<uint64_t>> part(pieces.size());
    const uint32_t *wv = g->walk_vtx;
    shard_parallel(pieces.size(), S / stride, [&](size_t i) {
        std::vector<uint64_t> &hist = part[i];
        hist.assign(BINS, 0);
        const uint64_t first = (pieces[i].s0 + stride - 1) / stride * stride;               // the sample is every stride-th step of the whole step array
        for (uint64_t s = first; s < pieces[i].s1; s += stride) hist[bin_of[wv[s]]]++;
    });
    std::vector<uint64_t> hist(BINS + 1, 0);
    for (const std::vector<uint64_t> &h : part) for (int b = 0; b < BINS; ++b) hist[b] += h[b];
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
    if (!walks_follow_order(g, coord)) return PHI_ERR_UNSUPPORTED;                           // no region cut for such walks
    slice_region(g, coord, k, w, coord_lo, coord_hi, slice_first, slice_len);
    return PHI_OK;
}

extern "C" int phi_shard_slice_walks_all(const phi_graph_view *g, int k, int w, int world, const uint64_t *coord_bounds,
                                         uint64_t *slice_first, uint64_t *slice_len)
{
    if (!g || !coord_bounds || !slice_first || !slice_len || k < 1 || w < 1 || world < 1) return PHI_ERR_ARG;
    std::vector<uint64_t> coord;
    if (!topo_coordinates(g, coord)) return PHI_ERR_UNSUPPORTED;
    if (!walks_follow_order(g, coord)) return PHI_ERR_UNSUPPORTED;                           // checked once for all regions
    for (int r = 0; r < world; ++r)
        slice_region(g, coord, k, w, coord_bounds[r], coord_bounds[r + 1], slice_first + (size_t)r * g->n_walks, slice_len + (size_t)r * g->n_walks);
    return PHI_OK;
}
