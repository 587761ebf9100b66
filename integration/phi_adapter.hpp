// phi_adapter.hpp — what a PHI maintainer adds to bind the reference to libphi_gpu_index.so.
//
// Header-only C++11, includes nothing from this repo except the C ABI header.  It is compiled INTO
// the reference's ILP_index.cpp (after `#include "ILP_index.h"`), and seam.inc replaces
// /root/reference/src/ILP_index.cpp:543-743 (from `std::vector<int32_t> hap_sizes(num_walks);` to
// the end of the "Filtered/Retained Minimizers" fprintf).  Everything after the seam — the Gurobi
// model construction (:757-1409), solve and back-trace — is untouched and consumes the same
// `Anchor_hits[i][j][k]` / `count_sp_r` it always did.
//
// INTEGRATION.md walks through the patch; integration/build_patched.py applies it to a scratch copy
// and builds oracle/_ref/PHI_gpu, which tests/test_gpu_dropin.py compares byte for byte against the
// unmodified reference's model dump.
#ifndef PHI_ADAPTER_HPP
#define PHI_ADAPTER_HPP

#include "phi_gpu_index.h"

#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#ifdef PHI_ADAPTER_TESTHOOK
#include "phi_adapter_testhook.hpp"      // test builds only: the result can come from a file (CPU tests of phi_model.hpp)
#endif

namespace phi_adapter {


// Replaces ILP_function lines 543-743 up to the result: runs the library and prints the same stderr lines
// (:556, :563, :611, :641, :724-735, :738-743) from the returned counters.  The caller frees the result with release().
inline const phi_index_result *run_front_end_result(ILP_index &ix, std::vector<std::pair<std::string, std::string> > &ip_reads, int32_t &count_sp_r)
{
    // ---- flat views of the members read_gfa() filled (ILP_index.cpp:20-155)
    std::vector<uint64_t> seg_off(1, 0), walk_off(1, 0), read_off(1, 0);
    std::string seg_bases, read_bases;
    std::vector<uint32_t> walk_vtx;
    for (uint32_t v = 0; v < ix.n_vtx; ++v) { seg_bases += ix.node_seq[v]; seg_off.push_back(seg_bases.size()); }
    for (uint32_t h = 0; h < ix.num_walks; ++h) {
        walk_vtx.insert(walk_vtx.end(), ix.paths[h].begin(), ix.paths[h].end());
        walk_off.push_back(walk_vtx.size());
    }
    for (size_t r = 0; r < ip_reads.size(); ++r) { read_bases += ip_reads[r].second; read_off.push_back(read_bases.size()); }

    phi_graph_view g;
    g.n_vtx = ix.n_vtx; g.seg_off = seg_off.data(); g.seg_bases = (const uint8_t *)seg_bases.data();
    g.n_walks = ix.num_walks; g.walk_off = walk_off.data(); g.walk_vtx = walk_vtx.data();
    g.top_order_map = ix.top_order_map.data();
    phi_reads_view rd;
    rd.n_reads = ip_reads.size(); rd.read_off = read_off.data(); rd.read_bases = (const uint8_t *)read_bases.data();
    phi_index_params prm;
    prm.k = ix.k_mer; prm.w = ix.window; prm.threshold = ix.threshold; prm.debug = ix.debug ? 1 : 0;

    const char *dev_env = getenv("PHI_GPU_DEVICE");
    phi_gpu_index_ctx *ctx = 0;
    phi_index_result *res = 0;
#ifdef PHI_ADAPTER_TESTHOOK
    if (const char *f = getenv("PHI_ADAPTER_RESULT_FILE")) res = const_cast<phi_index_result *>(load_result_file(f));   // CPU tests of the model block
#endif
    if (!res) {
        int rc = phi_gpu_index_create(dev_env ? atoi(dev_env) : -1, &ctx);
        if (rc == PHI_OK) rc = phi_gpu_index_run(ctx, &g, &rd, &prm, &res);
        if (rc != PHI_OK) {                   // the reference's error style: message on stderr, exit(1) (:105-106)
            fprintf(stderr, "Error: GPU ILP_index front end failed (%d): %s\n", rc, phi_gpu_last_error(ctx));
            exit(1);
        }
    }

    // ---- the log lines downstream scripts scrape (data/postprocessing_*.py)
    const double t = realtime() - mg_realtime0;
    std::cerr << "Number of Minimizers" << std::endl;                                              // :556
    for (uint32_t h = 0; h < ix.num_walks; ++h)
        fprintf(stderr, "%s : %d\n", ix.hap_id2name[h].c_str(), (int)res->minimizers_per_walk[h]);   // :563
    if (ix.debug && res->shared_kmer_hist) {                                                         // :593-606
        fprintf(stderr, "Shared fraction of unique kmers by haplotypes\n");
        for (uint32_t i = 1; i <= ix.num_walks; ++i)
            fprintf(stderr, "[Haplotypes: %d, fraction of unique shared kmers: %.5f]\n", (int)i,
                    (float)(int32_t)res->shared_kmer_hist[i] / (float)(int32_t)res->n_walk_kmers);
    }
    fprintf(stderr, "[M::%s::%.3f*%.2f] Haplotypes sketched\n", "ILP_function", t, cputime() / t);   // :611
    fprintf(stderr, "[M::%s::%.3f*%.2f] Indexed reads with spectrum size: %d\n", "ILP_function", t, cputime() / t, res->count_sp_r);  // :641

    count_sp_r = res->count_sp_r;
    std::cerr << "Number of Anchors" << std::endl;                                                 // :724
    for (uint32_t h = 0; h < ix.num_walks; ++h)
        fprintf(stderr, "%s : %d\n", ix.hap_id2name[h].c_str(), (int)res->anchors_per_walk[h]);      // :734
    const int64_t filtered_kmers = res->n_filtered, retained_kmers = count_sp_r - filtered_kmers;    // :719-721
    fprintf(stderr, "[M::%s::%.3f*%.2f] Filtered/Retained Minimizers: %.2f/%.2f%%\n", "ILP_function",
            realtime() - mg_realtime0, cputime() / (realtime() - mg_realtime0),
            (float)filtered_kmers / (float)count_sp_r * 100, (float)retained_kmers / (float)count_sp_r * 100);  // :738-743

    if (ctx) phi_gpu_index_destroy(ctx);          // the result outlives its ctx (phi_gpu_index_result_free is safe afterwards)
    return res;
}

inline void release(const phi_index_result *res)
{
#ifdef PHI_ADAPTER_TESTHOOK
    if (getenv("PHI_ADAPTER_RESULT_FILE")) return;                       // owned by the test hook
#endif
    phi_gpu_index_result_free(const_cast<phi_index_result *>(res));
}

inline int32_t member_walk(const phi_index_result *res, uint64_t m)
{
    return res->member_walk16 ? (int32_t)res->member_walk16[m] : res->member_walk32[m];
}

// Rebuilds the nested vectors the model construction indexes (:643, :716).  The result is the filter's own map (:680-709):
// per rank the groups in key order, per group one vertex list and its walks.
inline void fill_anchor_hits(const phi_index_result *res, uint32_t num_walks,
                             std::vector<std::vector<std::vector<std::vector<int32_t> > > > &Anchor_hits)
{
    Anchor_hits.assign(res->count_sp_r, std::vector<std::vector<std::vector<int32_t> > >(num_walks));
    const int32_t *vtx = res->group_vtx;
    for (int32_t r = 0; r < res->count_sp_r; ++r)
        for (uint32_t g = res->rank_off[r]; g < res->rank_off[r + 1]; ++g) {
            const std::vector<int32_t> list(vtx, vtx + res->group_len[g]);
            for (uint32_t m = res->group_member_off[g]; m < res->group_member_off[g + 1]; ++m)
                Anchor_hits[r][member_walk(res, m)].push_back(list);
            vtx += res->group_len[g];
        }
}

// The drop-in call of seam.inc: front end + nested vectors, exactly what ILP_function lines 543-743 leave behind.
inline void run_front_end(ILP_index &ix, std::vector<std::pair<std::string, std::string> > &ip_reads,
                          std::vector<std::vector<std::vector<std::vector<int32_t> > > > &Anchor_hits, int32_t &count_sp_r)
{
    const phi_index_result *res = run_front_end_result(ix, ip_reads, count_sp_r);
    fill_anchor_hits(res, ix.num_walks, Anchor_hits);
    release(res);
}

}  // namespace phi_adapter
#endif
