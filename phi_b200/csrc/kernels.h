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
    CTR_NONMONO = 12,      // some walk is not strictly increasing in top_order_map
    CTR_GROUPS = 13,       // distinct (rank, vertex list) groups in the filter table
    CTR_GROUP_OVERFLOW = 14,
    CTR_BAD_TOPO = 15,            // top_order_map is not an index into [0, n_vtx): chunk coordinates fall back to segment offsets
    CTR_ACTIVE_CHUNKS = 16,       // chunks that own at least one valid window
    CTR_UNIQUE_WINDOWS = 17,      // window end positions owned by representative chunks (what the walk kernel really sketches)
    CTR_DEDUPE_MISMATCH = 18,
    CTR_PATH_HITS = 19,
    CTR_CHUNK_FLAGS = 20,         // chunk-start flags set by the step pass (guards the chunk field of the packed scan)
    CTR_SEG_TOO_LONG = 21,        // a segment of 2^31 bases or more
    CTR_WALK_KMERS = 22,          // distinct walk-minimizer hashes (-d1 statistic)
    CTR_SURV_VTX = 23,            // vertices of all instantiated surviving anchors           // hits of all walks (every member chunk counts what its representative found)     // a chunk differs from its fingerprint representative (128-bit collision): rerun without sharing
    CTR_OUT_GROUPS = 24,          // surviving (rank, vertex list) groups of the result
    CTR_PATH_POS = 25,            // k-mer positions of the walks that lie in owned chunks (all of them unless a walk region is set)
    CTR_BAD_VTX = 26,             // a walk step names a vertex id >= n_vtx (caller error: PHI_ERR_ARG)
    CTR_COUNT = 27
};

enum { WALK_MODE_PROBE = 0, WALK_MODE_ALL = 1 };

// threads of a sketch tile CTA.  Walk tiles: 4 warps (small barrier domains, 6 CTAs per SM, chunk-sized tiles fill better);
// read tiles: 8 warps (the reads are one long coordinate system: bigger tiles amortise the per-tile set-up).
constexpr int WALK_TILE_THREADS = 128, READ_TILE_THREADS = 256;
constexpr int TILE_WINDOWS_PER_THREAD = 8;   // general path: window end positions per tile = 8 x threads
constexpr int SEG_PER_TILE = TILE_WINDOWS_PER_THREAD;   // a walk tile emits its hits in batches of one run per thread: at most this many contiguous hit segments

// Tile geometry as a function of w (sketch_tile.cuh).  For 9 <= w <= 65 a tile without non-ACGT bytes runs the
// register-resident core: every lane owns 8 consecutive k-mer positions, a warp 256, of which the first
// halo_lanes(w) lanes only feed the windows of the lanes behind them; the tile is padded in front so that the halo
// window (the one before the tile's first own window) is the first window of warp 0.
__host__ __device__ inline bool tile_fast_w(int w) { return w >= 9 && w <= 65; }
__host__ __device__ inline int tile_halo_lanes(int w) { return (w + 6) >> 3; }                      // ceil((w - 1) / 8)
__host__ __device__ inline int tile_pad(int w) { return tile_fast_w(w) ? 8 * tile_halo_lanes(w) - (w - 1) : 0; }
__host__ __device__ inline int tile_cap(int w, int threads) { return tile_fast_w(w) ? (threads / 32) * (32 - tile_halo_lanes(w)) * 8 - 1 : TILE_WINDOWS_PER_THREAD * threads; }

// One walk-sketch tile: window end positions [e0, e1) of walk `walk` (the representative of chunk `chunk`).
struct TileRec {
    uint32_t walk, e0, e1, first_step;   // first_step: step under the tile's first staged base, relative to walk_off[walk]
    uint32_t chunk, cbase, walk_len, _r1; // cbase: first base of the chunk; hit positions are stored as p + w - cbase; walk_len: bases of the walk
    uint64_t step0, step_end;            // global index of that first step and of the walk's last step + 1: the sketch kernel needs no
};                                       // walk_off / walk_len lookups of its own (one dependent round trip less per tile)

// Chunk table (chunks.cu): chunks are numbered along the walks, walk after walk.
struct ChunkTable {
    uint32_t n_chunks;
    uint32_t *chunk_step;                // [n_chunks + 1] first step (global index) of each chunk
    uint32_t *c_walk, *c_L, *c_R;        // walk; context steps [L, R] (global indices)
    uint32_t *c_lo, *c_hi;               // owned window end positions [lo, hi) in walk coordinates (lo == hi: nothing to sketch)
    uint64_t *c_h1, *c_h2;               // fingerprint of the context
    uint32_t *c_slot, *c_rep;            // grouping table slot; representative chunk (0xFFFFFFFF: inactive)
    uint32_t *c_ninst;                   // members of the group (at the representative)
    uint32_t *c_ntile, *c_tile_base;     // tiles of a representative and their first index
};

struct ExpandArgs {
    int w; uint32_t walk_id_base;
    const uint64_t *member_off;          // [n_chunks + 1] first expanded record of each member chunk
    const uint32_t *hseg_off, *hseg_cnt; // [n_tiles * SEG_PER_TILE]
    const uint8_t *rank_drop;
    const uint32_t *hit_rank, *hit_pos; const uint64_t *hit_voff; const uint8_t *hit_nv; const uint64_t *hit_hash;
    uint32_t *x_rank, *x_walk, *x_pos; uint64_t *x_voff; uint8_t *x_nv; uint64_t *x_hash;
};

// shared-memory carve-up of a sketch tile (computed on the host: sketch_tile.cuh make_layout)
struct TileLayout {
    int M, M8, NB, nchunks, pad, cap;   // pad: positions in front of the halo window; cap: own windows per tile
    int o_canon, o_hash, o_pack, o_dirty, o_bnd, o_first, o_scan, o_pre, o_suf, o_flag, o_base, o_stepv, o_steps, o_cfirst, o_cmask;
    int o_raw;                          // read tiles: landing zone of the bulk copy (cp.async.bulk) of the tile's bytes
    int bytes;
};
TileLayout read_tile_layout(int k, int w);
TileLayout walk_tile_layout(int k, int w);

struct ReadSketchArgs {
    TileLayout layout;
    const uint8_t *read_bases;        // 16 readable bytes in front, >= 16 zero bytes after total_bases
    const uint64_t *read_off;         // [n_reads + 1]
    uint64_t n_reads, total_bases;
    const uint64_t *tile_first_read;  // [n_tiles]
    uint64_t tile0;                   // first tile of this launch (the reads may be sketched piece by piece while they arrive)
    int k, w;
    uint64_t *table; uint64_t table_mult, table_limit;   // home slot = umulhi(key, table_mult); slots [0, table_limit), table[table_limit] stays EMPTY
    int bulk;                         // stage the tile's bytes with one cp.async.bulk into shared memory (issued before the boundary pass) instead of per-thread loads
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
    const TileRec *tiles;                                     // [n_tiles] tiles of the representative chunks
    int k, w, mode;
    const uint64_t *spec; const uint32_t *dir; int dbits;     // ranked spectrum + radix directory
    uint32_t *chunk_emitted, *chunk_hits;                     // [n_chunks] minimizers emitted / hits found by a representative chunk
    uint32_t *hseg_off, *hseg_cnt;                            // [n_tiles * SEG_PER_TILE] hit segments of each tile, in position order
    // hits of the representative chunks: hit_chunk = chunk id, hit_pos = p + w - (first base of the chunk)
    uint32_t *hit_rank, *hit_chunk, *hit_pos; uint64_t *hit_voff; uint8_t *hit_nv; uint64_t *hit_hash;
    int32_t *vtx_pool;
    uint4 *probe;                                             // optional [2 * hit_cap]: the group table's 32-byte probe record of every hit (filter.cu)
    uint64_t hit_cap, vtx_cap;
    unsigned long long *ctr;
};

int read_tile_windows(int w);

// sketch_kernels.cu
cudaError_t launch_read_tile_dir(const uint64_t *read_off, uint64_t n_reads, int w, uint64_t n_tiles, uint64_t *out, cudaStream_t st);
cudaError_t launch_read_sketch(const ReadSketchArgs &A, uint64_t n_tiles, cudaStream_t st);
cudaError_t launch_walk_sketch(const WalkSketchArgs &A, uint32_t n_tiles, cudaStream_t st);
cudaError_t launch_hash_bytes(const uint8_t *keys, uint64_t n, int len, uint64_t *out, cudaStream_t st);

// One u32 per walk step: bits 0..30 = bases of the step's segment, bit 31 = the step starts a chunk.  Scanned as a u64 with the
// chunk flag moved up to bit STEP_BASE_BITS: one scan yields (chunks before the step, bases before the step).
constexpr int STEP_BASE_BITS = 38;       // < 2^38 walk bases and < 2^26 chunks per GPU (checked on the host)
struct PackedStep {
    uint32_t v;
    __host__ __device__ operator uint64_t() const { return (uint64_t)(v & 0x7FFFFFFFu) | ((uint64_t)(v >> 31) << STEP_BASE_BITS); }
};

// chunks.cu — walk preparation (step lengths and bases, walk lengths), walk chunking, grouping of identical chunks,
// instantiation of the representatives' hits
// per-vertex record of the step pass: (bases, coordinate >> shift, top_order_map, region) with region = 0 / 1 / 2 for a
// coordinate below / inside / at or above the owned range [own_lo, own_hi) (phi_gpu_index_set_walk_region; everything is owned by default)
cudaError_t chunk_topo_coord(const int32_t *top_order_map, const uint64_t *seg_off, uint32_t n_vtx, int shift, uint64_t own_lo, uint64_t own_hi,
                             uint32_t *tlen, uint64_t *prefix, uint4 *vinfo, void *scan_scratch, unsigned long long *ctr, cudaStream_t st, uint64_t *launches);
// per step: segment length, chunk-start flag, zero-length count, topological monotonicity -> packed[]
cudaError_t walk_step_pass(const uint32_t *walk_vtx, const uint64_t *walk_off, uint32_t n_walks, uint64_t n_steps, const uint4 *vinfo, uint32_t n_vtx,
                           PackedStep *packed, unsigned long long *ctr, cudaStream_t st, uint64_t *launches);
// step pass + scan + finalisation in one kernel (decoupled look-back; the common case without zero-length steps, which it only
// counts): step_base, walk_len, chunk_step[0..chunks] / c_walk (both sized for n_steps + 1 chunks), ctr[CTR_CHUNK_FLAGS] = chunks.
// tile_state: walk_steps_fused_tiles(n_steps) words.
// tile_first / tile_last: the tiles [tile_first, tile_last) are launched (the steps may arrive piece by piece: tiles take their index
// from a ticket, so consecutive launches continue the scan); tile_first == 0 also resets the scan state.
uint64_t walk_steps_fused_tiles(uint64_t n_steps);
uint64_t walk_steps_fused_tile_steps();
cudaError_t walk_steps_fused(const uint32_t *walk_vtx, const uint64_t *walk_off, uint32_t n_walks, uint64_t n_steps, const uint4 *vinfo, uint32_t n_vtx,
                             unsigned long long *tile_state, uint32_t *ticket, uint32_t *step_base, uint32_t *chunk_step, uint32_t *c_walk,
                             uint64_t *walk_len, unsigned long long *ctr, cudaStream_t st, uint64_t *launches,
                             uint64_t tile_first = 0, uint64_t tile_last = ~0ull);
// scanned = exclusive scan of packed (as u64) -> step_base, walk_len, chunk_step / c_walk of C
cudaError_t walk_step_finalize(const ChunkTable &C, const PackedStep *packed, const uint64_t *scanned, const uint64_t *walk_off, uint32_t n_walks,
                               uint64_t n_steps, uint32_t *step_base, uint64_t *walk_len, cudaStream_t st, uint64_t *launches);
// drop the zero-length steps (rare: segments without bases): out_vtx / out_off describe the compacted walks
cudaError_t walk_compact_steps(const uint32_t *walk_vtx, const uint64_t *walk_off, uint32_t n_walks, uint64_t n_steps, const PackedStep *packed,
                               uint32_t *flags, uint64_t *pos, void *scan_scratch, uint64_t n_kept, uint32_t *out_vtx, uint64_t *out_off,
                               cudaStream_t st, uint64_t *launches);
cudaError_t chunk_keys(const ChunkTable &C, const uint32_t *walk_vtx, const uint64_t *walk_off, const uint32_t *step_base, const uint64_t *walk_len,
                       const uint4 *vinfo, int k, int w, unsigned long long *ctr, cudaStream_t st, uint64_t *launches);
cudaError_t chunk_group(const ChunkTable &C, uint32_t *table, uint32_t table_cap, const uint32_t *walk_vtx, int dedupe, int w, unsigned long long *ctr,
                        cudaStream_t st, uint64_t *launches);
cudaError_t chunk_tiles(const ChunkTable &C, const uint64_t *walk_off, const uint64_t *walk_len, const uint32_t *step_base, int w, TileRec *tiles, cudaStream_t st, uint64_t *launches);
cudaError_t chunk_emitted(const ChunkTable &C, const uint32_t *c_emitted, const uint32_t *c_hits, uint32_t walk_id_base,
                          unsigned long long *minimizers_per_walk, unsigned long long *ctr, cudaStream_t st, uint64_t *launches);
cudaError_t chunk_survivors(const ChunkTable &C, const TileRec *tiles, uint32_t n_tiles, const uint32_t *hseg_off, const uint32_t *hseg_cnt,
                            const uint32_t *hit_rank, const uint8_t *hit_nv, const uint8_t *rank_drop, uint32_t *c_surv, uint32_t *c_surv_vtx,
                            uint32_t *member_cnt, unsigned long long *ctr, cudaStream_t st, uint64_t *launches);
cudaError_t chunk_expand(const ChunkTable &C, const ExpandArgs &X, cudaStream_t st, uint64_t *launches);


// primitives.cu — all on `st`, scratch supplied by the caller
// exclusive scan of u32 -> u64 (out may not alias in); returns bytes of scratch needed when scratch == nullptr
size_t scan_u32_to_u64_scratch(uint64_t n);
cudaError_t scan_u32_to_u64(const uint32_t *in, uint64_t *out, uint64_t n, void *scratch, cudaStream_t st, uint64_t *launches);
// exclusive scan of u32 (n < 2^32 total), in place or out of place
size_t scan_u32_scratch(uint64_t n);
cudaError_t scan_u32_inplace(uint32_t *data, uint64_t n, void *scratch, cudaStream_t st, uint64_t *launches);
cudaError_t scan_u32(const uint32_t *in, uint32_t *out, uint64_t n, void *scratch, cudaStream_t st, uint64_t *launches);
// exclusive scan of packed walk steps as u64 (scratch: scan_u32_to_u64_scratch)
cudaError_t scan_packed_steps(const PackedStep *in, uint64_t *out, uint64_t n, void *scratch, cudaStream_t st, uint64_t *launches);
// stable LSD radix sort of u64 keys (optional u32 values) on bits [bit_lo, bit_hi); result ends in keys_a/vals_a
size_t radix_sort_scratch(uint64_t n);
cudaError_t radix_sort_u64(uint64_t *keys_a, uint64_t *keys_b, uint32_t *vals_a, uint32_t *vals_b, uint64_t n, int bit_lo, int bit_hi,
                           void *scratch, cudaStream_t st, uint64_t *launches);
// stable LSD radix sort of (u32 key, u32 value) records on the low `bits` bits, digits of up to 11 bits; result in keys_a/vals_a
size_t radix_sort_u32_scratch(uint64_t n);
cudaError_t radix_sort_u32(uint32_t *keys_a, uint32_t *keys_b, uint32_t *vals_a, uint32_t *vals_b, uint64_t n, int bits,
                           void *scratch, cudaStream_t st, uint64_t *launches);
// Order-preserving spectrum table (sketch_kernels.cu: table_insert): slots [0, limit) + one EMPTY sentinel at table[limit].
constexpr uint64_t TABLE_PAD = 4096;      // probe room behind the last home slot (no wrap-around)
constexpr uint64_t TABLE_BLOCK = 2048;    // slots per block of the count / write kernels
// sort every probe cluster in place -> the occupied slots are in ascending order; then count per block
size_t table_blocks(uint64_t limit);
cudaError_t table_insert_keys(const uint64_t *keys, uint64_t n, uint64_t *table, uint64_t base, uint64_t mult, uint64_t limit,
                              unsigned long long *ctr, cudaStream_t st, uint64_t *launches);
cudaError_t table_sort_and_count(uint64_t *table, uint64_t limit, uint32_t *block_cnt, cudaStream_t st, uint64_t *launches);
// block_off = exclusive scan of block_cnt: write the occupied slots in slot order to out
cudaError_t table_write_ordered(const uint64_t *table, uint64_t limit, const uint32_t *block_off, uint64_t *out, cudaStream_t st, uint64_t *launches);
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
    int rank_bits;                   // bits of the largest rank (sort width)
};
struct FilterWork {                  // device scratch, sized by the host
    uint2 *g_slot; uint64_t g_cap;                 // group table: slot = (representing record, count), one 8-byte word
    const uint4 *probe;                            // [2 * n_hits] one 32-byte probe record per hit: (rank, vertices, first 6 vertices)
    uint32_t *hit_slot;                            // [n_hits] slot of each hit's group
    uint32_t *hit_sub;                             // optional [n_hits]: the group's count before this hit was added (its sub-offset in the group)
    const uint32_t *weight;                        // optional [n_hits]: occurrences a record stands for (multi-GPU summaries)
    const uint32_t *chunk_weight;                  // optional [n_chunks]: members of the chunk hit_walk[i] names (hits of representatives)
    uint8_t *rank_drop;                            // [n_ranks] (4-byte aligned, padded to a multiple of 4)
    int mark_inline;                               // the counts of this table are global: apply the threshold while counting
    uint32_t *flags;                               // [n_hits] survivor flags -> scanned
    uint64_t *keys_a, *keys_b; uint32_t *vals_a, *vals_b;   // [n_survivors]
    void *sort_scratch; void *scan_scratch;
    unsigned long long *ctr;
};
cudaError_t filter_shared_kmer_hist(const uint64_t *hash, const uint32_t *walk, uint64_t n, uint32_t n_walks, unsigned long long *hist,
                                    unsigned long long *distinct, cudaStream_t st, uint64_t *launches);
// probe records of the hits of A (filter_count_groups compares a hit with the record that represents a slot through them: one
// 32-byte sector instead of a chain of four dependent scattered reads)
cudaError_t filter_build_probe(const FilterArgs &A, uint4 *probe, cudaStream_t st, uint64_t *launches);
cudaError_t filter_count_groups(const FilterArgs &A, const FilterWork &W, cudaStream_t st, uint64_t *launches);
// stable sort of the records (which arrive in (walk, position) order) on their rank: order ends up in W.vals_a
cudaError_t filter_sort_records(const FilterArgs &A, const FilterWork &W, cudaStream_t st, uint64_t *launches);
// by_walk != 0: order[] is sorted by (rank, walk, position) and (rank, walk) runs are re-ordered; by_walk == 0: order[] holds one
// record per group sorted by rank and the runs of one rank are re-ordered (all keys distinct there)
cudaError_t filter_fix_multi(const FilterArgs &A, uint32_t *order, uint64_t n_surv, int by_walk, uint32_t *big_list, uint32_t big_cap,
                             unsigned long long *ctr, cudaStream_t st, uint64_t *launches);
// big (rank, walk) groups recorded by filter_fix_multi (count in ctr[CTR_BIG_GROUPS], read on the device: no host round trip)
cudaError_t filter_fix_big(const FilterArgs &A, uint32_t *order, uint32_t *tmp, const uint32_t *big_list, uint32_t big_cap,
                           const unsigned long long *ctr, cudaStream_t st, uint64_t *launches);
cudaError_t filter_csr_sizes(const FilterArgs &A, const uint32_t *order, uint64_t n_surv, uint32_t *nv_out, uint8_t *anchor_len, cudaStream_t st, uint64_t *launches);
cudaError_t filter_csr_fill(const FilterArgs &A, const uint32_t *order, uint64_t n_surv, const uint64_t *anchor_off, uint64_t *rank_off,
                            int32_t *anchor_walk, int32_t *anchor_vtx, unsigned long long *anchors_per_walk,
                            uint32_t n_walks_out, cudaStream_t st, uint64_t *launches);

// groups.cu — the grouped result: member walks per representative chunk, surviving groups, their members and vertex lists
constexpr uint32_t SORT_VALS_MAX = 2048;   // walks per GPU up to which member lists are merged by a shared-memory counting sort
constexpr uint32_t GROUP_HIST_MAX = 4096;  // walks up to which the anchors-per-walk histogram is block-local
struct GroupOut {
    const uint32_t *order;                         // [n_groups] representing hit of every group, final order
    const uint32_t *member_off, *vtx_off;          // [n_groups + 1]
    const uint32_t *cm_off, *cm_walk;              // member walks (local ids, ascending) of every representative chunk
    uint32_t *members_tmp;                         // [n_members] parts as the hits wrote them
    void *member_walk; int member_walk_bytes;      // result: u16 walk ids when they all fit (member_walk_bytes == 2), else i32
    int32_t *group_vtx;
    unsigned long long *anchors_per_walk;
    uint32_t walk_id_base;
};
// cm_off = exclusive scan of c_ninst; cm_walk = walks of the members of every representative, ascending
cudaError_t chunk_members(const ChunkTable &C, uint32_t n_walks, uint32_t *cm_off, uint32_t *cursor, uint32_t *cm_tmp, uint32_t *cm_walk,
                          void *scan_scratch, cudaStream_t st, uint64_t *launches);
// flags/pos: [n_hits] scratch; keys/vals receive (rank, representing hit) of every surviving group, in hit order;
// ctr[CTR_OUT_GROUPS / CTR_SURVIVORS / CTR_SURV_VTX] += groups / members / vertices
cudaError_t groups_compact(const FilterArgs &A, const FilterWork &W, uint32_t *flags, uint32_t *pos, uint32_t *keys, uint32_t *vals,
                           void *scan_scratch, cudaStream_t st, uint64_t *launches);
cudaError_t groups_sizes(const FilterArgs &A, const FilterWork &W, const uint32_t *order, uint32_t n_groups, uint32_t *cnt_out, uint32_t *nv_out,
                         uint8_t *group_len, uint32_t *rank_off, cudaStream_t st, uint64_t *launches);
cudaError_t groups_fill(const FilterArgs &A, const FilterWork &W, const GroupOut &G, uint32_t n_groups, uint32_t n_walks_local, uint32_t n_walks_out,
                        cudaStream_t st, uint64_t *launches);

}  // namespace phi
