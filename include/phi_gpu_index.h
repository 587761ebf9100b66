/*
 * phi_gpu_index.h — C ABI of the B200-native PHI front end (ILP_index stage).
 *
 * The reference has no plugin/FFI interface; the seam this library replaces is
 * INSIDE ILP_index::ILP_function, between /root/reference/src/ILP_index.cpp:543
 * (first statement after the "Graph has ..." log line) and :752 (end of the
 * in_nodes loop is :745-752 and stays on the caller's side; the replaced range
 * is :543-743).  Caller before the seam: main() (/root/reference/src/main.cpp:140).
 * Consumer after the seam: the Gurobi model construction reading
 * Anchor_hits[i][j][k] and count_sp_r (/root/reference/src/ILP_index.cpp:786-833,
 * :838-879).  INTEGRATION.md shows the adapter a PHI maintainer would add.
 *
 * Plain pointers and sizes only: no C++ types, no exceptions, no torch types
 * cross this boundary.  There is NO CPU fallback: every entry point that
 * computes fails with PHI_ERR_CUDA when no sm_100 device is usable.
 *
 * Ownership: the caller owns every input buffer for the duration of a call;
 * the library owns a phi_index_result until phi_gpu_index_result_free().
 * A ctx is single-threaded (one ctx per host thread / per GPU), re-entrant
 * across ctxs.
 */
#ifndef PHI_GPU_INDEX_H
#define PHI_GPU_INDEX_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHI_GPU_INDEX_ABI_VERSION 3

/* status codes (0 == ok); text via phi_gpu_last_error() */
enum {
    PHI_OK = 0,
    PHI_ERR_ARG = 1,          /* bad argument / inconsistent view */
    PHI_ERR_UNSUPPORTED = 2,  /* legal for the reference but not implemented here (e.g. k > 255, w > 256): hard error, never a silent divergence */
    PHI_ERR_CUDA = 3,         /* CUDA runtime / no device */
    PHI_ERR_NOMEM = 4,
    PHI_ERR_COMM = 5          /* NCCL / multi-GPU exchange */
};

/*
 * Graph, flat view of what ILP_index::read_gfa() builds
 * (/root/reference/src/ILP_index.cpp:20-155):
 *   seg_off/seg_bases  <- node_seq[v]   (:35)  original case, ASCII, concatenated
 *   walk_off/walk_vtx  <- paths[h]      (:110) forward-strand vertex ids, W-line order
 *   top_order_map      <- top_order_map (:150-154) Kahn order index per vertex
 * Vertex id = segment index by first appearance (/root/reference/src/gfa-base.cpp:75-96).
 */
typedef struct {
    uint32_t n_vtx;
    const uint64_t *seg_off;        /* [n_vtx + 1] */
    const uint8_t *seg_bases;       /* [seg_off[n_vtx]] */
    uint32_t n_walks;
    const uint64_t *walk_off;       /* [n_walks + 1] */
    const uint32_t *walk_vtx;       /* [walk_off[n_walks]] */
    const int32_t *top_order_map;   /* [n_vtx] */
} phi_graph_view;

/* Reads, flat view of ip_reads[r].second (/root/reference/src/ILP_index.cpp:313-328, :620). */
typedef struct {
    uint64_t n_reads;
    const uint64_t *read_off;       /* [n_reads + 1] */
    const uint8_t *read_bases;      /* [read_off[n_reads]] ASCII, any case */
} phi_reads_view;

/*
 * Parameters = the ILP_index members the replaced range reads
 * (/root/reference/src/ILP_index.h:72-81; set in /root/reference/src/main.cpp:118-131):
 * k_mer (-k, 31), window (-w, 25), threshold (-T, 1.0f), debug (-d).
 * Accepted: 1 <= k <= 255 (k-mers of up to 32 bases are packed two bits per base; longer ones are compared and
 * hashed from their spelling, like k-mers with non-ACGT bytes), 1 <= w <= 256.
 */
typedef struct {
    int32_t k;
    int32_t w;
    float threshold;
    int32_t debug;
} phi_index_params;

/*
 * Result = everything that crosses the seam.
 *
 * count_sp_r / spectrum : ILP_index.cpp:631-636 — number of distinct read
 *     minimizer hashes and the hashes themselves in ascending unsigned order;
 *     rank r <-> spectrum[r].
 * groups (FINAL post-filter order) : ILP_index.cpp:670-716.  The reference's filter builds, per rank, a std::map from the
 *     key string of a vertex list to its members (walk, anchor) and re-emits the anchors group after group in key order.
 *     The result keeps exactly that form: the groups of rank r are [rank_off[r], rank_off[r + 1]) in the reference's key
 *     order; group g has group_len[g] vertices (1..k), the lists lie back to back in group_vtx in group order, and its
 *     members are the walks member_walk[group_member_off[g] .. group_member_off[g + 1]) in ascending order (a walk occurs
 *     twice when it carries the list twice); member_walk is member_walk16 (u16 walk ids, half the bytes over PCIe) when
 *     n_walks <= 65536, else member_walk32.  One forward pass
 *         for r: for g in groups(r): for h in members(g): Anchor_hits[r][h].push_back(list(g))
 *     rebuilds the reference's nested vectors exactly (integration/phi_adapter.hpp) — it is the loop at :700-709.
 *     n_anchors = number of members = anchors in Anchor_hits.  The form is compact on purpose: the result crosses PCIe, and
 *     with h haplotypes the per-anchor form would repeat every vertex list up to h times.
 *     Sketch-only results (phi_gpu_index_sketch_walks): one group per emitted minimizer in (walk, path position) order,
 *     rank_off == NULL, group_member_off == NULL (group g has the single member member_walk32[g]).
 * minimizers_per_walk : kmer_index[h].size(), log line ILP_index.cpp:563.
 * anchors_per_walk    : log lines ILP_index.cpp:725-735.
 * n_filtered          : filtered_kmers, ILP_index.cpp:719-721 (log :738-743).
 * n_walk_kmers / shared_kmer_hist : the -d1 statistic, ILP_index.cpp:565-606 — number of distinct walk-minimizer hashes
 *     (uniqe_kmers.size()) and, for i in [0, n_walks], how many of them occur in exactly i walks (kmer_hist_count[i]).
 *     Only computed when params.debug != 0 (shared_kmer_hist == NULL otherwise); with several GPUs every rank returns the same, global statistic.
 */
typedef struct {
    int32_t count_sp_r;
    uint32_t n_walks;
    int64_t n_filtered;
    uint64_t n_anchors;                  /* members over all groups */
    uint64_t n_groups;
    uint64_t n_group_vtx;
    const uint64_t *spectrum;            /* [count_sp_r] */
    const uint32_t *rank_off;            /* [count_sp_r + 1] first group of every rank; NULL for sketch-only results */
    const uint8_t *group_len;            /* [n_groups] */
    const int32_t *group_vtx;            /* [n_group_vtx] */
    const uint32_t *group_member_off;    /* [n_groups + 1]; NULL for sketch-only results (one member per group) */
    const uint16_t *member_walk16;       /* [n_anchors] when n_walks <= 65536 (and not sketch-only), else NULL */
    const int32_t *member_walk32;        /* [n_anchors] otherwise; exactly one of the two is non-NULL */
    const uint64_t *minimizers_per_walk; /* [n_walks] */
    const uint64_t *anchors_per_walk;    /* [n_walks] */
    /* work counters of this run (for throughput / roofline arithmetic) */
    uint64_t read_kmer_positions;        /* sum_r max(0, len_r - k + 1) over reads with len_r >= w + k - 1 */
    uint64_t path_kmer_positions;        /* same over walks */
    uint64_t read_minimizers_emitted;    /* table inserts attempted */
    uint64_t path_minimizers_emitted;    /* sum of minimizers_per_walk (== probes) */
    uint64_t path_hits;                  /* probes that found a rank (pre-filter anchors) */
    /* debug statistic (params.debug) */
    uint64_t n_walk_kmers;
    const uint64_t *shared_kmer_hist;    /* [n_walks + 1] or NULL */
} phi_index_result;

/* Per-stage device times of the last run, CUDA events on the ctx's two streams (ms).  The graph preparation (second stream)
 * runs concurrently with the read stage (main stream), so the stage times add up to more than total_ms. */
typedef struct {
    float h2d_ms;            /* first event -> last host -> device copy done (0 for a resident run); overlaps the read stage */
    float graph_prep_ms;     /* second stream: step bases, walk chunks, fingerprints, grouping, tiles */
    float read_sketch_ms;    /* read minimizers -> order-preserving HBM hash table -> sorted distinct hashes (waits for H2D of the reads) */
    float spectrum_ms;       /* radix directory (multi-GPU: + exchange of the spectrum) */
    float walk_sketch_ms;    /* minimizers of the representative chunks + probe + anchor emit */
    float filter_ms;         /* group count, threshold, group order, member walks and vertex lists of the surviving groups */
    float d2h_ms;            /* device -> host copies of the result */
    float total_ms;          /* first event to last event */
    float walk_kernel_ms;    /* the walk sketch kernel alone (roofline numerator's denominator) */
    float read_kernel_ms;    /* the read sketch kernel alone */
    uint64_t kernel_launches;/* kernels launched by this library during the run */
    /* multi-GPU runs only (0 otherwise): parts of spectrum_ms and filter_ms spent in the NCCL exchanges */
    float exchange_spectrum_ms; /* hash-range all-to-all + owner dedup/sort + slice broadcast */
    float route_hits_ms;        /* local group table + one summary per group */
    float exchange_hits_ms;     /* all-to-all of the group summaries + owner-side counts + broadcast of the drop flags */
} phi_stage_times;

typedef struct phi_gpu_index_ctx phi_gpu_index_ctx;

/* device < 0 selects the current device.  Fails (PHI_ERR_CUDA) without a GPU. */
int phi_gpu_index_create(int device, phi_gpu_index_ctx **out);
void phi_gpu_index_destroy(phi_gpu_index_ctx *ctx);
/* Last error text of this ctx, or with ctx == NULL of the last failed create / comm_unique_id ON THE CALLING THREAD (the text is
 * kept per thread: fetch it on the thread that made the call). Never NULL. */
const char *phi_gpu_last_error(const phi_gpu_index_ctx *ctx);
int phi_gpu_index_abi_version(void);

/*
 * The drop-in call: replaces ILP_index.cpp:543-743.  Host buffers in, host
 * result out; H2D and D2H copies happen inside.
 */
int phi_gpu_index_run(phi_gpu_index_ctx *ctx, const phi_graph_view *graph, const phi_reads_view *reads,
                      const phi_index_params *params, phi_index_result **out);

/*
 * Resident variant (used to time the device path alone): upload once, then run
 * any number of times on the HBM-resident inputs.  `download` != 0 copies the
 * result to the host (then *out is set); with download == 0 *out receives a
 * result whose array pointers are NULL but whose scalar counters are valid.
 */
int phi_gpu_index_upload(phi_gpu_index_ctx *ctx, const phi_graph_view *graph, const phi_reads_view *reads);
int phi_gpu_index_run_resident(phi_gpu_index_ctx *ctx, const phi_index_params *params, int download,
                               phi_index_result **out);

/* Result arrays live in pinned host memory borrowed from the ctx; _result_free hands it back for reuse by the next
 * run (free a result before the next run on the same ctx and no new pinning happens).  Safe after ctx destroy. */
void phi_gpu_index_result_free(phi_index_result *res);
/* Pinned (page-locked) host memory for INPUT buffers: views that point into it are copied to the device at full
 * PCIe speed; ordinary pageable memory works too, only slower. */
void *phi_gpu_host_alloc(size_t bytes);
void phi_gpu_host_free(void *p);
int phi_gpu_index_last_times(const phi_gpu_index_ctx *ctx, phi_stage_times *out);

/*
 * Walk sharing (no counterpart in the reference, which sketches every walk on its own,
 * /root/reference/src/ILP_index.cpp:556-573): walks are cut into chunks at content-defined boundaries, chunks with the
 * same vertex sequence and context are sketched ONCE, and the hits are instantiated for every walk that contains the
 * chunk.  The result is bit-identical with or without sharing; these knobs only change how much work is done.
 *   chunk_shift : log2 of the chunk bucket size in bases of the topological coordinate (default 11; 4..24)
 *   share       : 0 = sketch every chunk of every walk on its own, 1 = share identical chunks (default)
 */
int phi_gpu_index_set_walk_sharing(phi_gpu_index_ctx *ctx, int chunk_shift, int share);
typedef struct {
    uint64_t chunks;            /* chunks over all walks of this ctx */
    uint64_t active_chunks;     /* chunks that own at least one window */
    uint64_t tiles;             /* sketch tiles launched (tiles of the representative chunks) */
    uint64_t unique_windows;    /* window end positions actually sketched (<= path_kmer_positions) */
    uint64_t unique_hits;       /* hits found by the representatives (<= path_hits) */
} phi_walk_sharing_stats;
int phi_gpu_index_last_sharing(const phi_gpu_index_ctx *ctx, phi_walk_sharing_stats *out);

/*
 * Walk sketch alone == ILP_index::index_kmers for every walk
 * (/root/reference/src/ILP_index.cpp:359-445): every emitted minimizer of every
 * walk, in path order, with its hash and its vertex list.  Returned through a
 * phi_index_result in which  spectrum == NULL,  rank_off == NULL (count_sp_r == 0),
 * anchors are in (walk, path position) order and  *hashes_out[a]  is the
 * minimizer hash.  Free *hashes_out with phi_gpu_index_free_u64.
 */
int phi_gpu_index_sketch_walks(phi_gpu_index_ctx *ctx, const phi_graph_view *graph, const phi_index_params *params,
                               phi_index_result **out, uint64_t **hashes_out);
void phi_gpu_index_free_u64(uint64_t *p);

/*
 * Device MurmurHash3_x64_128(key, len, seed 0) -> out[0]^out[1], i.e. the
 * reference's hash128_to_64 (/root/reference/src/ILP_index.cpp:10-18;
 * /root/reference/src/MurmurHash3.cpp:255-332), over n keys of `len` bytes
 * each, packed back to back.  Known-answer-test hook.
 */
int phi_gpu_hash128_to_64(phi_gpu_index_ctx *ctx, const uint8_t *keys, uint64_t n, int32_t len, uint64_t *out);

/*
 * Multi-GPU (one ctx per GPU, one process or thread per ctx).  The 128-byte
 * id comes from phi_gpu_index_comm_unique_id() on rank 0 and is distributed by
 * the caller (torch.distributed store, MPI, a file ...).  After comm_init,
 * phi_gpu_index_run / _run_resident treat `reads` and the walks of `graph` as
 * THIS RANK'S SHARD (segments and top_order_map replicated): whole walks
 * (walk_id_base = id of the first one) or, with phi_gpu_index_set_walk_region,
 * all walks cut to this rank's region of the graph (walk_id_base 0).  The ranks
 * exchange read minimizer hashes by hash range (all-to-all), all-gather the
 * owners' sorted slices, route one (rank, count, vertex list) summary per local
 * group to the owner of the rank (which applies the threshold; the drop flags
 * are all-reduced), and every rank returns the surviving groups of ITS walks /
 * region for all ranks (spectrum: rank 0 only, NULL elsewhere - it is the same
 * everywhere; n_filtered: the dropped ranks this rank owns; sum over the ranks);
 * phi_index_result_merge makes one result of the parts.  params.debug and every
 * k <= 255 work as on one GPU.  A stage that fails on one rank is reported to
 * all ranks through the next exchange (PHI_ERR_COMM on the others); a failure in
 * the middle of an exchange aborts the communicator (comm_init again to go on).
 * phi_gpu_index_destroy and a repeated comm_init release the communicator.
 * walk ids in the result are global: walk_id_base + local index.
 */
#define PHI_COMM_ID_BYTES 128
int phi_gpu_index_comm_unique_id(uint8_t id[PHI_COMM_ID_BYTES]);
int phi_gpu_index_comm_init(phi_gpu_index_ctx *ctx, int rank, int world, const uint8_t id[PHI_COMM_ID_BYTES],
                            uint32_t walk_id_base, uint32_t n_walks_global);

/*
 * Host-side ingest (no GPU needed), written from scratch: what gfa_read + ILP_index::read_gfa
 * (/root/reference/src/gfa-io.cpp:462-508, /root/reference/src/ILP_index.cpp:20-155) and kseq + read_ip_reads
 * (/root/reference/src/kseq.h:192-232, /root/reference/src/ILP_index.cpp:313-328) produce, straight into the flat
 * views above (.gz or plain).  Vertex ids, walk order, walk names and read order are the reference's; top_order_map is a
 * Kahn order of the same graph (ties between equally valid orders may be broken differently, results cannot differ).
 * Errors: PHI_ERR_ARG (cannot open), PHI_ERR_UNSUPPORTED (a walk with reverse-strand steps, as ILP_index.cpp:104-107);
 * the message goes to err[errlen] when given.
 */
typedef struct phi_host_graph phi_host_graph;
typedef struct phi_host_reads phi_host_reads;
int phi_host_graph_load(const char *gfa_path, phi_host_graph **out, char *err, size_t errlen);
const phi_graph_view *phi_host_graph_view(const phi_host_graph *g);
const char *phi_host_graph_walk_name(const phi_host_graph *g, uint32_t walk);       /* sample + "." + haplotype index */
const char *phi_host_graph_segment_name(const phi_host_graph *g, uint32_t vtx);
uint64_t phi_host_graph_n_links(const phi_host_graph *g);
/* walk steps (u -> v consecutive in a walk) that no L-line backs.  0 for graphs written by pangenome builders; only then is the order
 * of the vertices inside an anchor independent of how ties between equally valid topological orders are broken (the reference's
 * Kahn order, /root/reference/src/ILP_index.cpp:116-147, and this loader's may break them differently). */
uint64_t phi_host_graph_unlinked_steps(const phi_host_graph *g);
void phi_host_graph_free(phi_host_graph *g);
int phi_host_reads_load(const char *path, phi_host_reads **out, char *err, size_t errlen);
const phi_reads_view *phi_host_reads_view(const phi_host_reads *r);
const char *phi_host_reads_name(const phi_host_reads *r, uint64_t i);
void phi_host_reads_free(phi_host_reads *r);

/*
 * Region partition of the walks over several GPUs (no counterpart in the reference: one process, OpenMP over whole walks,
 * /root/reference/src/ILP_index.cpp:559-562).  Splitting BY WALK would undo walk sharing, so every GPU gets ALL walks cut to the
 * steps whose vertices lie in its range [coord_lo, coord_hi) of the topological base coordinate (bases of all vertices that
 * precede a vertex in top_order_map) plus context (>= w bases in front, >= k-1 bases behind):
 *   phi_shard_walk_regions   : world+1 coordinate bounds, balanced by walk steps (the walks of `g` may be a sample);
 *   phi_shard_slice_walks    : for every walk of `g` the slice [slice_first[h], slice_first[h] + slice_len[h]) of g->walk_vtx
 *                              (absolute step indices) that the GPU owning [coord_lo, coord_hi) needs; the caller builds that GPU's
 *                              graph view from the slices (same walk order, walk_id_base 0, n_walks_global = n_walks);
 *   phi_shard_slice_walks_all : the same for all `world` regions of coord_bounds at once (slice_first / slice_len are [world][n_walks],
 *                              region r at r * n_walks): the O(steps) check that the walks follow the order runs once, on several
 *                              host threads, instead of once per region;
 *   phi_gpu_index_set_walk_region : tells the ctx which coordinate range it owns: of the walks (slices) it is given, only the
 *                              windows whose last k-mer starts on a vertex inside the range are produced (path_kmer_positions,
 *                              minimizers_per_walk and the groups then cover the owned part only; sum / merge over the GPUs).
 *                              Default: [0, 2^64) — everything.
 *   phi_index_result_merge   : the per-GPU parts -> ONE result in the reference's order: groups of a rank merged in key order
 *                              (/root/reference/src/ILP_index.cpp:680-709), member lists of a group that occurs in several parts
 *                              united, per-walk counters / n_filtered / work counters summed, spectrum from the part that has it.
 *                              Works for parts of a by-walk partition (walk_id_base) as well.  Free with phi_gpu_index_result_free.
 *                              Big merges cut the hash ranks into blocks and run them on up to 16 host threads
 *                              (PHI_MERGE_THREADS overrides; count pass, prefix sum, write pass), output arrays are 2 MB aligned
 *                              and marked for transparent huge pages.
 * PHI_ERR_UNSUPPORTED: top_order_map is not a permutation, or a walk does not follow it (use the by-walk partition then).
 */
int phi_shard_walk_regions(const phi_graph_view *g, int world, uint64_t *coord_bounds /* [world + 1] */);
int phi_shard_slice_walks(const phi_graph_view *g, int k, int w, uint64_t coord_lo, uint64_t coord_hi,
                          uint64_t *slice_first /* [n_walks] */, uint64_t *slice_len /* [n_walks] */);
int phi_shard_slice_walks_all(const phi_graph_view *g, int k, int w, int world, const uint64_t *coord_bounds /* [world + 1] */,
                              uint64_t *slice_first /* [world * n_walks] */, uint64_t *slice_len /* [world * n_walks] */);
int phi_gpu_index_set_walk_region(phi_gpu_index_ctx *ctx, uint64_t coord_lo, uint64_t coord_hi);
int phi_index_result_merge(const phi_index_result *const *parts, int n_parts, phi_index_result **out);

/* Host-only partition helpers (no GPU needed). */
/* owner rank of a hash under the range partition on the high bits */
int phi_shard_owner_of_hash(uint64_t hash, int world);
/* contiguous, size-balanced split of n items with per-item weights off[i+1]-off[i] into `world` parts:
 * writes world+1 boundaries into bounds[] */
int phi_shard_split_by_weight(const uint64_t *off, uint64_t n, int world, uint64_t *bounds);

#ifdef __cplusplus
}
#endif
#endif /* PHI_GPU_INDEX_H */
