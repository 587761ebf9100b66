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
#include "phi_shards.hpp"     // detail::Shard, detail::build_shards: the per-GPU views of a multi-GPU run

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#ifdef PHI_ADAPTER_TESTHOOK
#include "phi_adapter_testhook.hpp"      // test builds only: the result can come from a file (CPU tests of phi_model.hpp)
#endif

namespace phi_adapter {


// What the front end returns: ONE result in the reference's order (the per-GPU parts of a multi-GPU run are merged by
// phi_index_result_merge).  `parts` stays a list for the test hook, which can feed by-walk parts from files: walk ranges ascend
// with p there, so "the parts one after another" is the reference's walk order.
struct FrontEnd {
    std::vector<const phi_index_result *> parts;
    bool from_files;
    FrontEnd() : from_files(false) {}
    int32_t count_sp_r() const { return parts.empty() ? 0 : parts[0]->count_sp_r; }
};

inline int32_t member_walk(const phi_index_result *res, uint64_t m)
{
    return res->member_walk16 ? (int32_t)res->member_walk16[m] : res->member_walk32[m];
}

namespace detail {
struct Job { phi_gpu_index_ctx *ctx; int rank, world; const uint8_t *id; uint32_t n_walks_global; Shard *sh; const phi_index_params *prm; phi_index_result *res; int rc; std::string err; };
inline void run_job(Job *j)
{
    phi_gpu_index_ctx *ctx = j->ctx;
    j->rc = PHI_OK;
    if (j->world > 1) j->rc = phi_gpu_index_comm_init(ctx, j->rank, j->world, j->id, j->sh->walk_id_base, j->n_walks_global);
    if (j->rc == PHI_OK) j->rc = j->sh->region ? phi_gpu_index_set_walk_region(ctx, j->sh->coord_lo, j->sh->coord_hi) : phi_gpu_index_set_walk_region(ctx, 0, ~0ull);
    if (j->rc == PHI_OK) j->rc = phi_gpu_index_run(ctx, &j->sh->g, &j->sh->rd, j->prm, &j->res);
    if (j->rc != PHI_OK) j->err = phi_gpu_last_error(ctx);
}

// The ctxs of this process: one per GPU named by PHI_GPU_DEVICES=a,b,c (or PHI_GPU_DEVICE=n, default: the current device).  They are
// created ONCE, by a thread that starts when the program is loaded, so that CUDA context creation (a few hundred ms) runs under the
// reference's own GFA / read parsing instead of inside the front end, and they are kept for every later call (all device buffers
// and the pinned result pool are grow-only members of a ctx: a second call allocates nothing).
struct Pool {
    std::vector<int> devices;
    std::vector<phi_gpu_index_ctx *> ctx;
    std::vector<int> rc; std::vector<std::string> err;
    std::thread th; bool started, joined;
    Pool() : started(false), joined(false) {}
    ~Pool() { if (started && !joined && th.joinable()) th.join(); }      // (the ctxs themselves live until the process ends)
    void start()
    {
        if (started) return;
        started = true;
        if (const char *dl = getenv("PHI_GPU_DEVICES")) {
            std::string list(dl);
            for (size_t p = 0; p < list.size();) { size_t c = list.find(',', p); if (c == std::string::npos) c = list.size(); devices.push_back(atoi(list.substr(p, c - p).c_str())); p = c + 1; }
        }
        if (devices.empty()) { const char *d = getenv("PHI_GPU_DEVICE"); devices.push_back(d ? atoi(d) : -1); }
        ctx.assign(devices.size(), (phi_gpu_index_ctx *)0); rc.assign(devices.size(), PHI_OK); err.assign(devices.size(), std::string());
        th = std::thread([this]() {
            std::vector<std::thread> each;
            for (size_t i = 0; i < devices.size(); ++i)
                each.push_back(std::thread([this, i]() { rc[i] = phi_gpu_index_create(devices[i], &ctx[i]); if (rc[i] != PHI_OK) err[i] = phi_gpu_last_error(0); }));
            for (size_t i = 0; i < each.size(); ++i) each[i].join();
        });
    }
    void wait() { start(); if (!joined) { th.join(); joined = true; } }
};
inline Pool &pool() { static Pool p; return p; }
struct Warm { Warm() { if (!getenv("PHI_ADAPTER_RESULT_FILE")) pool().start(); } };
static Warm g_warm;                                // program load: start creating the ctxs
}  // namespace detail

// Replaces ILP_function lines 543-743 up to the result: runs the library and prints the same stderr lines
// (:556, :563, :611, :641, :724-735, :738-743) from the returned counters.  The caller frees the result with release().
// PHI_GPU_DEVICE=n selects the GPU; PHI_GPU_DEVICES=a,b,c runs one ctx per listed GPU, one host thread each: reads sharded by bases,
// walks by REGION of the topological coordinate (phi_shard_walk_regions / phi_shard_slice_walks: walk sharing keeps working across
// GPUs; by whole walks when the graph does not allow a region cut), NCCL inside the library, parts merged by phi_index_result_merge.
inline FrontEnd run_front_end_result(ILP_index &ix, std::vector<std::pair<std::string, std::string> > &ip_reads, int32_t &count_sp_r)
{
    phi_index_params prm;
    prm.k = ix.k_mer; prm.w = ix.window; prm.threshold = ix.threshold; prm.debug = ix.debug ? 1 : 0;

    FrontEnd fe;
#ifdef PHI_ADAPTER_TESTHOOK
    if (const char *f = getenv("PHI_ADAPTER_RESULT_FILE")) {            // CPU tests of the model blocks: "a.bin" or "a.bin,b.bin,..." (parts)
        fe.from_files = true;
        std::string list(f);
        for (size_t p = 0; p <= list.size();) {
            size_t c = list.find(',', p); if (c == std::string::npos) c = list.size();
            fe.parts.push_back(load_result_file(list.substr(p, c - p).c_str()));
            p = c + 1;
        }
    }
#endif
    if (fe.parts.empty()) {
        const double t_enter = realtime();
        // ---- flat views of the members read_gfa() filled (ILP_index.cpp:20-155): sized once, filled in parallel
        std::vector<uint64_t> seg_off(ix.n_vtx + 1, 0), walk_off(ix.num_walks + 1, 0), read_off(ip_reads.size() + 1, 0);
        for (uint32_t v = 0; v < ix.n_vtx; ++v) seg_off[v + 1] = seg_off[v] + ix.node_seq[v].size();
        for (uint32_t h = 0; h < ix.num_walks; ++h) walk_off[h + 1] = walk_off[h] + ix.paths[h].size();
        for (size_t r = 0; r < ip_reads.size(); ++r) read_off[r + 1] = read_off[r] + ip_reads[r].second.size();
        std::string seg_bases(seg_off[ix.n_vtx], '\0'), read_bases(read_off[ip_reads.size()], '\0');
        std::vector<uint32_t> walk_vtx(walk_off[ix.num_walks]);
        #pragma omp parallel for schedule(static) num_threads(ix.num_threads > 0 ? ix.num_threads : 1)
        for (int64_t v = 0; v < (int64_t)ix.n_vtx; ++v) ix.node_seq[v].copy(&seg_bases[seg_off[v]], ix.node_seq[v].size());
        #pragma omp parallel for schedule(dynamic, 1) num_threads(ix.num_threads > 0 ? ix.num_threads : 1)
        for (int64_t h = 0; h < (int64_t)ix.num_walks; ++h) std::copy(ix.paths[h].begin(), ix.paths[h].end(), walk_vtx.begin() + walk_off[h]);
        #pragma omp parallel for schedule(static) num_threads(ix.num_threads > 0 ? ix.num_threads : 1)
        for (int64_t r = 0; r < (int64_t)ip_reads.size(); ++r) ip_reads[r].second.copy(&read_bases[read_off[r]], ip_reads[r].second.size());

        const double t_flat = realtime();
        detail::Pool &P = detail::pool();
        P.wait();
        const double t_ctx = realtime();
        const int W = (int)P.devices.size();
        for (int r = 0; r < W; ++r)
            if (P.rc[r] != PHI_OK) { fprintf(stderr, "Error: GPU ILP_index front end failed (%d): %s\n", P.rc[r], P.err[r].c_str()); exit(1); }
        phi_graph_view gfull;
        gfull.n_vtx = ix.n_vtx; gfull.seg_off = seg_off.data(); gfull.seg_bases = (const uint8_t *)seg_bases.data();
        gfull.n_walks = ix.num_walks; gfull.walk_off = walk_off.data(); gfull.walk_vtx = walk_vtx.data(); gfull.top_order_map = ix.top_order_map.data();
        phi_reads_view rfull;
        rfull.n_reads = ip_reads.size(); rfull.read_off = read_off.data(); rfull.read_bases = (const uint8_t *)read_bases.data();
        uint8_t id[PHI_COMM_ID_BYTES];
        if (W > 1 && phi_gpu_index_comm_unique_id(id) != PHI_OK) { fprintf(stderr, "Error: %s\n", phi_gpu_last_error(0)); exit(1); }
        std::vector<detail::Shard> shards;
        detail::build_shards(gfull, rfull, W, prm.k, prm.w, shards);
        const double t_shard = realtime();
        std::vector<detail::Job> jobs(W);
        for (int r = 0; r < W; ++r) {
            detail::Job &j = jobs[r];
            j.ctx = P.ctx[r]; j.rank = r; j.world = W; j.id = id; j.n_walks_global = ix.num_walks; j.sh = &shards[r]; j.prm = &prm; j.res = 0; j.rc = PHI_OK;
        }
        if (W == 1) detail::run_job(&jobs[0]);
        else {
            std::vector<std::thread> th;
            for (int r = 0; r < W; ++r) th.push_back(std::thread(detail::run_job, &jobs[r]));
            for (int r = 0; r < W; ++r) th[r].join();
        }
        std::vector<const phi_index_result *> parts;
        for (int r = 0; r < W; ++r) {
            if (jobs[r].rc != PHI_OK) {           // the reference's error style: message on stderr, exit(1) (:105-106)
                fprintf(stderr, "Error: GPU ILP_index front end failed (%d): %s\n", jobs[r].rc, jobs[r].err.c_str());
                exit(1);
            }
            parts.push_back(jobs[r].res);
        }
        if (getenv("PHI_ADAPTER_TIMES"))          // where the front end's wall time goes (stderr, one line)
            fprintf(stderr, "[phi_adapter] flat views %.4f s, wait for the CUDA context(s) %.4f s, shards %.4f s, run on %d GPU(s) %.4f s\n",
                    t_flat - t_enter, t_ctx - t_flat, t_shard - t_ctx, W, realtime() - t_shard);
        if (W == 1) fe.parts = parts;
        else {                                    // one result in the reference's order
            phi_index_result *merged = 0;
            const int mrc = phi_index_result_merge(parts.data(), W, &merged);
            if (mrc != PHI_OK) { fprintf(stderr, "Error: merging the per-GPU results failed (%d)\n", mrc); exit(1); }
            for (int r = 0; r < W; ++r) phi_gpu_index_result_free(const_cast<phi_index_result *>(parts[r]));
            fe.parts.push_back(merged);
        }
    }

    // ---- the log lines downstream scripts scrape (data/postprocessing_*.py); per-walk counters and n_filtered are partial sums
    const phi_index_result *res = fe.parts[0];
    const double t = realtime() - mg_realtime0;
    std::cerr << "Number of Minimizers" << std::endl;                                              // :556
    for (uint32_t h = 0; h < ix.num_walks; ++h) {
        uint64_t n = 0; for (size_t p = 0; p < fe.parts.size(); ++p) n += fe.parts[p]->minimizers_per_walk[h];
        fprintf(stderr, "%s : %d\n", ix.hap_id2name[h].c_str(), (int)n);                            // :563
    }
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
    for (uint32_t h = 0; h < ix.num_walks; ++h) {
        uint64_t n = 0; for (size_t p = 0; p < fe.parts.size(); ++p) n += fe.parts[p]->anchors_per_walk[h];
        fprintf(stderr, "%s : %d\n", ix.hap_id2name[h].c_str(), (int)n);                            // :734
    }
    int64_t filtered_kmers = 0; for (size_t p = 0; p < fe.parts.size(); ++p) filtered_kmers += fe.parts[p]->n_filtered;
    const int64_t retained_kmers = count_sp_r - filtered_kmers;                                      // :719-721
    fprintf(stderr, "[M::%s::%.3f*%.2f] Filtered/Retained Minimizers: %.2f/%.2f%%\n", "ILP_function",
            realtime() - mg_realtime0, cputime() / (realtime() - mg_realtime0),
            (float)filtered_kmers / (float)count_sp_r * 100, (float)retained_kmers / (float)count_sp_r * 100);  // :738-743
    return fe;
}

inline void release(FrontEnd &fe)
{
    if (!fe.from_files)                                                    // (file-fed parts are owned by the test hook)
        for (size_t p = 0; p < fe.parts.size(); ++p) phi_gpu_index_result_free(const_cast<phi_index_result *>(fe.parts[p]));
    fe.parts.clear();
}

// Rebuilds the nested vectors the model construction indexes (:643, :716).  A result is the filter's own map (:680-709):
// per rank the groups in key order, per group one vertex list and its walks.
inline void fill_anchor_hits(const FrontEnd &fe, uint32_t num_walks,
                             std::vector<std::vector<std::vector<std::vector<int32_t> > > > &Anchor_hits)
{
    Anchor_hits.assign(fe.count_sp_r(), std::vector<std::vector<std::vector<int32_t> > >(num_walks));
    for (size_t p = 0; p < fe.parts.size(); ++p) {
        const phi_index_result *res = fe.parts[p];
        const int32_t *vtx = res->group_vtx;
        for (int32_t r = 0; r < res->count_sp_r; ++r)
            for (uint32_t g = res->rank_off[r]; g < res->rank_off[r + 1]; ++g) {
                const std::vector<int32_t> list(vtx, vtx + res->group_len[g]);
                for (uint32_t m = res->group_member_off[g]; m < res->group_member_off[g + 1]; ++m)
                    Anchor_hits[r][member_walk(res, m)].push_back(list);
                vtx += res->group_len[g];
            }
    }
}

// The drop-in call of seam.inc: front end + nested vectors, exactly what ILP_function lines 543-743 leave behind.
inline void run_front_end(ILP_index &ix, std::vector<std::pair<std::string, std::string> > &ip_reads,
                          std::vector<std::vector<std::vector<std::vector<int32_t> > > > &Anchor_hits, int32_t &count_sp_r)
{
    FrontEnd fe = run_front_end_result(ix, ip_reads, count_sp_r);
    fill_anchor_hits(fe, ix.num_walks, Anchor_hits);
    release(fe);
}

}  // namespace phi_adapter
#endif
