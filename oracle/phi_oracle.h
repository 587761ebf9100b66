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
 * Uses the same plain-C view/result structs as the product ABI so results can
 * be compared field by field.
 */
#ifndef PHI_ORACLE_H
#define PHI_ORACLE_H
#include "../include/phi_gpu_index.h"
#ifdef __cplusplus
extern "C" {
#endif

/* hash128_to_64: ILP_index.cpp:10-18 over MurmurHash3_x64_128 (seed 0). */
uint64_t phi_oracle_hash128_to_64(const uint8_t *key, int32_t len);
void phi_oracle_murmur3_x64_128(const uint8_t *key, int32_t len, uint32_t seed, uint64_t out[2]);

/* ILP_function lines 543-743 on flat views.  n_threads <= 0: all cores. */
int phi_oracle_index_run(const phi_graph_view *graph, const phi_reads_view *reads, const phi_index_params *params,
                         int n_threads, phi_index_result **out);

/* index_kmers for every walk (ILP_index.cpp:359-445); layout as phi_gpu_index_sketch_walks. */
int phi_oracle_sketch_walks(const phi_graph_view *graph, const phi_index_params *params, int n_threads,
                            phi_index_result **out, uint64_t **hashes_out);

/* compute_hashes for one read (ILP_index.cpp:447-493): sorted distinct hashes; returns count, fills *out (malloc). */
int64_t phi_oracle_read_hashes(const uint8_t *read, uint64_t len, int32_t k, int32_t w, uint64_t **out);

void phi_oracle_result_free(phi_index_result *res);
void phi_oracle_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
