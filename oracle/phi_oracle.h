/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of PHI's ILP_index
 * front end, /root/reference/src/ILP_index.cpp:10-18, 330-357, 359-445,
 * 447-493, 495-526, 543-743 and /root/reference/src/MurmurHash3.cpp:255-332.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this code against
 * (a) the nine MurmurHash3 known-answer vectors produced by the reference's own
 * MurmurHash3.cpp, (b) per-walk minimizer lists and per-read hash sets dumped by
 * oracle/_ref/ref_probe (the unmodified reference's index_kmers/compute_hashes),
 * (c) the anchors parsed back out of the model dump of oracle/_ref/PHI_ref (the
 * unmodified reference CLI), all committed under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call
 * into this library.  The product (phi_b200/) never does.
 *
 * Uses the product ABI's plain-C input views and parameters.  The result has its own struct: the reference's final
 * Anchor_hits flattened anchor by anchor in (rank, walk, j) order — what the reference-side adapter rebuilds from the
 * product's grouped result, so the two are compared after that expansion.
 */
#ifndef PHI_ORACLE_H
#define PHI_ORACLE_H
#include "../include/phi_gpu_index.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t count_sp_r;
    uint32_t n_walks;
    int64_t n_filtered;
    uint64_t n_anchors;
    uint64_t n_anchor_vtx;
    const uint64_t *spectrum;            /* [count_sp_r] ascending */
    const uint64_t *rank_off;            /* [count_sp_r + 1] first anchor of every rank; NULL for sketch-only results */
    const int32_t *anchor_walk;          /* [n_anchors] anchor a is Anchor_hits[rank][anchor_walk[a]][j] */
    const uint8_t *anchor_len;           /* [n_anchors] vertices of anchor a */
    const int32_t *anchor_vtx;           /* [n_anchor_vtx] the lists back to back */
    const uint64_t *minimizers_per_walk; /* [n_walks] */
    const uint64_t *anchors_per_walk;    /* [n_walks] */
    uint64_t read_kmer_positions, path_kmer_positions, read_minimizers_emitted, path_minimizers_emitted, path_hits;
    uint64_t n_walk_kmers;
    const uint64_t *shared_kmer_hist;    /* [n_walks + 1] with params.debug, else NULL */
} phi_oracle_result;

/* hash128_to_64: ILP_index.cpp:10-18 over MurmurHash3_x64_128 (seed 0). */
uint64_t phi_oracle_hash128_to_64(const uint8_t *key, int32_t len);
void phi_oracle_murmur3_x64_128(const uint8_t *key, int32_t len, uint32_t seed, uint64_t out[2]);

/* ILP_function lines 543-743 on flat views.  n_threads <= 0: all cores. */
int phi_oracle_index_run(const phi_graph_view *graph, const phi_reads_view *reads, const phi_index_params *params,
                         int n_threads, phi_oracle_result **out);

/* index_kmers for every walk (ILP_index.cpp:359-445): one anchor per emitted minimizer in (walk, path position) order,
 * rank_off == NULL; (*hashes_out)[a] = the minimizer hash. */
int phi_oracle_sketch_walks(const phi_graph_view *graph, const phi_index_params *params, int n_threads,
                            phi_oracle_result **out, uint64_t **hashes_out);

/* The same, plus for every emitted minimizer the vertex under the start of the LAST k-mer of the window that emitted it (the
 * loop variable i of ILP_index.cpp:388 at :413): what the multi-GPU region partition decides ownership by.  Free with phi_oracle_free. */
int phi_oracle_sketch_walks_owner(const phi_graph_view *graph, const phi_index_params *params, int n_threads,
                                  phi_oracle_result **out, uint64_t **hashes_out, int32_t **owner_vtx_out);

/* compute_hashes for one read (ILP_index.cpp:447-493): sorted distinct hashes; returns count, fills *out (malloc). */
int64_t phi_oracle_read_hashes(const uint8_t *read, uint64_t len, int32_t k, int32_t w, uint64_t **out);

void phi_oracle_result_free(phi_oracle_result *res);
void phi_oracle_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
