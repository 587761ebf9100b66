// Internal declarations shared by the .cu translation units of libphi_gpu_index.so.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace phi {

constexpr uint64_t TABLE_EMPTY = 0xFFFFFFFFFFFFFFFFull;

// device counter block (unsigned long long each)
enum {
    CTR_DISTINCT = 0,      // distinct keys inserted into the spectrum table
    CTR_OVERFLOW = 1,      // table full
    CTR_HAS_MAXKEY = 2,    // the key ~0 occurred (kept outside the table)
    CTR_READ_EMITTED = 3,  // read minimizers emitted
    CTR_HITS = 4,          // walk hits
    CTR_HIT_VTX = 5,       // vertices of walk hits
    CTR_FILTERED = 6,      // ranks dropped by the threshold filter
    CTR_SURVIVORS = 7,
    CTR_BIG_GROUPS = 8,
    CTR_ZERO_STEPS = 9,
    CTR_READ_POS = 10,
    CTR_COMPACT = 11,
    CTR_NONMONO = 12,
    CTR_GROUPS = 13,       // distinct (rank, vertex list) groups in the filter table
    CTR_GROUP_OVERFLOW = 14,      // some walk is not strictly increasing in top_order_map
    CTR_COUNT = 16
};

enum { WALK_MODE_PROBE = 0, WALK_MODE_ALL = 1 };

// shared-memory carve-up of a sketch tile (computed on the host: sketch_tile.cuh make_layout)
struct TileLayout {
    int M, M8, NB, nchunks;
    int o_canon, o_hash, o_pack, o_dirty, o_bnd, o_scan, o_pre, o_suf, o_flag, o_base, o_stepv, o_steps, o_cfirst, o_cmask;
    int bytes;
};
TileLayout tile_layout(int k, int w, bool walk);

struct ReadSketchArgs {
    TileLayout layout;
    const uint8_t *read_bases;        // 16 readable bytes in front, >= 16 zero bytes after total_bases
    const uint64_t *read_off;         // [n_reads + 1]
    uint64_t n_reads, total_bases;
    const uint64_t *tile_first_read;  // [n_tiles]
    int k, w;
    uint64_t *table; uint64_t table_mask;
    unsigned long long *ctr;
};

struct WalkSketchArgs {
    TileLayout layout;
    int walks_monotone;                                       // every walk strictly increasing in top_order_map: anchors need no per-hit check
    const uint8_t *seg_bases;                                 // 16 readable bytes in front, >= 16 after
    const uint64_t *seg_off; const int32_t *top_order_map;
    const uint32_t *walk_vtx; const uint64_t *walk_off;      // zero-length steps removed
    const uint32_t *step_base;                                // walk-relative first base of each step
    const uint64_t *walk_len;                                 // [n_walks] bases
    const uint64_t *walk_tile_base;                           // [n_walks + 1]
    const uint32_t *tile_first_step;                          // [total tiles] relative to walk_off[h]
    int k, w, mode;
    const uint64_t *spec; const uint32_t *dir; int dbits;     // ranked spectrum + radix directory
    uint32_t walk_id_base;
    unsigned long long *minimizers_per_walk;                  // [n_walks]
    uint32_t *hit_rank, *hit_walk, *hit_pos; uint64_t *hit_voff; uint8_t *hit_nv; uint64_t *hit_hash;
    int32_t *vtx_pool;
    uint64_t hit_cap, vtx_cap;
    unsigned long long *ctr;
};

int tile_windows();

// sketch_kernels.cu
cudaError_t launch_read_tile_dir(const uint64_t *read_off, uint64_t n_reads, int w, uint64_t n_tiles, uint64_t *out, cudaStream_t st);
cudaError_t launch_read_sketch(const ReadSketchArgs &A, uint64_t n_tiles, cudaStream_t st);
cudaError_t launch_walk_sketch(const WalkSketchArgs &A, uint32_t n_walks, uint64_t max_tiles, cudaStream_t st);
cudaError_t launch_step_len(const uint32_t *walk_vtx, const uint64_t *seg_off, uint64_t n_steps, uint32_t *step_len, cudaStream_t st);
cudaError_t launch_walk_len(const uint64_t *gbase, const uint32_t *step_len, const uint64_t *walk_off, uint32_t n_walks,
                            uint64_t n_steps, uint64_t *walk_len, cudaStream_t st);
cudaError_t launch_step_finalize(const uint64_t *gbase, const uint32_t *step_len, const uint64_t *walk_off, uint32_t n_walks,
                                 uint64_t n_steps, int w, const uint64_t *walk_tile_base, uint32_t *step_base,
                                 uint32_t *tile_first_step, cudaStream_t st);
cudaError_t launch_walk_monotone(const uint32_t *walk_vtx, const uint64_t *walk_off, uint32_t n_walks, uint64_t n_steps,
                                 const int32_t *top_order_map, unsigned long long *ctr, cudaStream_t st);
cudaError_t launch_hash_bytes(const uint8_t *keys, uint64_t n, int len, uint64_t *out, cudaStream_t st);

// primitives.cu — all on `st`, scratch supplied by the caller
// exclusive scan of u32 -> u64 (out may not alias in); returns bytes of scratch needed when scratch == nullptr
size_t scan_u32_to_u64_scratch(uint64_t n);
cudaError_t scan_u32_to_u64(const uint32_t *in, uint64_t *out, uint64_t n, void *scratch, cudaStream_t st, uint64_t *launches);
// in-place exclusive scan of u32 (n < 2^32 total)
size_t scan_u32_scratch(uint64_t n);
cudaError_t scan_u32_inplace(uint32_t *data, uint64_t n, void *scratch, cudaStream_t st, uint64_t *launches);
// stable LSD radix sort of u64 keys (optional u32 values) on bits [bit_lo, bit_hi); result ends in keys_a/vals_a
size_t radix_sort_scratch(uint64_t n);
cudaError_t radix_sort_u64(uint64_t *keys_a, uint64_t *keys_b, uint32_t *vals_a, uint32_t *vals_b, uint64_t n, int bit_lo, int bit_hi,
                           void *scratch, cudaStream_t st, uint64_t *launches);
// compact the non-empty slots of the spectrum table into out (order arbitrary); count written to *d_count
cudaError_t table_compact(const uint64_t *table, uint64_t cap, uint64_t *out, unsigned long long *d_count, cudaStream_t st, uint64_t *launches);
// radix directory over the top dbits bits of a sorted key array: dir[b] = lower_bound(prefix b), dir[2^dbits] = n
cudaError_t build_directory(const uint64_t *sorted, uint32_t n, int dbits, uint32_t *dir, cudaStream_t st, uint64_t *launches);
cudaError_t fill_u64(uint64_t *p, uint64_t n, uint64_t v, cudaStream_t st, uint64_t *launches);
cudaError_t fill_u32(uint32_t *p, uint64_t n, uint32_t v, cudaStream_t st, uint64_t *launches);

// filter.cu
struct FilterArgs {
    uint64_t n_hits;
    const uint32_t *hit_rank, *hit_walk, *hit_pos; const uint64_t *hit_voff; const uint8_t *hit_nv; const int32_t *vtx_pool;
    uint32_t n_ranks;
    float thr;                       // threshold * num_walks, evaluated in float as the reference does (:698)
    const uint64_t *walk_gbase;      // [n_walks_global + 1] global base coordinate of each walk start (ordering key)
    int gpos_bits, rank_bits;
};
struct FilterWork {                  // device scratch, sized by the host
    uint32_t *g_rep, *g_cnt; uint64_t g_cap;      // group table
    uint32_t *hit_slot;                            // [n_hits] slot of each hit's group
    const uint32_t *weight;                        // optional [n_hits]: occurrences a record stands for (multi-GPU summaries)
    uint8_t *rank_drop;                            // [n_ranks]
    uint32_t *flags;                               // [n_hits] survivor flags -> scanned
    uint64_t *keys_a, *keys_b; uint32_t *vals_a, *vals_b;   // [n_survivors]
    void *sort_scratch; void *scan_scratch;
    unsigned long long *ctr;
};
cudaError_t filter_count_groups(const FilterArgs &A, const FilterWork &W, cudaStream_t st, uint64_t *launches);
cudaError_t filter_mark_drops(const FilterArgs &A, const FilterWork &W, cudaStream_t st, uint64_t *launches);
cudaError_t filter_flag_survivors(const FilterArgs &A, const FilterWork &W, cudaStream_t st, uint64_t *launches);
cudaError_t filter_emit_keys(const FilterArgs &A, const FilterWork &W, uint64_t n_surv, bool combined, cudaStream_t st, uint64_t *launches);
cudaError_t filter_fix_multi(const FilterArgs &A, uint32_t *order, uint64_t n_surv, uint32_t *big_list, uint32_t big_cap,
                             unsigned long long *ctr, cudaStream_t st, uint64_t *launches);
cudaError_t filter_fix_big(const FilterArgs &A, uint32_t *order, uint32_t *tmp, const uint32_t *big_list, uint32_t n_big,
                           uint64_t n_surv, cudaStream_t st, uint64_t *launches);
cudaError_t filter_csr_sizes(const FilterArgs &A, const uint32_t *order, uint64_t n_surv, uint32_t *nv_out, cudaStream_t st, uint64_t *launches);
cudaError_t filter_csr_fill(const FilterArgs &A, const uint32_t *order, uint64_t n_surv, const uint64_t *anchor_off, int32_t *anchor_rank,
                            int32_t *anchor_walk, int32_t *anchor_vtx, unsigned long long *anchors_per_walk, uint32_t walk_id_base,
                            uint32_t n_walks_out, cudaStream_t st, uint64_t *launches);

}  // namespace phi
