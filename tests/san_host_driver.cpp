// TEST INFRASTRUCTURE (CPU): the threaded host code of the library — the loaders (parallel BGZF inflate, parallel chunked read parse,
// streamed parse), the partition helpers and phi_index_result_merge — linked straight from their sources into one program that
// tests/test_sanitizers.py builds with -fsanitize=thread and -fsanitize=address,undefined.
//   san_host_driver load <file>...     load every file (GFA by name, else reads); graphs are also cut into 4 regions
//   san_host_driver merge              random result dealt to 5 parts member by member, merged on 1 / 3 / 8 threads, compared
#include "../include/phi_gpu_index.h"
#include "../phi_b200/csrc/result_box.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

extern "C" void phi_gpu_index_result_free(phi_index_result *res)      // the heap branch of the library's own (phi_gpu_index.cu is not linked here)
{
    if (!res) return;
    ResultBox *b = (ResultBox *)res;
    for (int i = 0; i < b->nbufs; ++i) free(b->bufs[i].p);
    free(b);
}

static int load(int n, char **files)
{
    char err[512];
    for (int i = 0; i < n; ++i) {
        const std::string s(files[i]);
        if (s.find(".gfa") == std::string::npos) {
            phi_host_reads *R = 0;
            const int rc = phi_host_reads_load(files[i], &R, err, sizeof err);
            printf("%s reads rc=%d n=%llu bases=%llu\n", files[i], rc, rc ? 0ull : (unsigned long long)phi_host_reads_view(R)->n_reads,
                   rc || !phi_host_reads_view(R)->n_reads ? 0ull : (unsigned long long)phi_host_reads_view(R)->read_off[phi_host_reads_view(R)->n_reads]);
            if (!rc) phi_host_reads_free(R);
        } else {
            phi_host_graph *G = 0;
            const int rc = phi_host_graph_load(files[i], &G, err, sizeof err);
            printf("%s graph rc=%d", files[i], rc);
            if (!rc) {
                const phi_graph_view *g = phi_host_graph_view(G);
                std::vector<uint64_t> b(5), f(4 * (size_t)g->n_walks + 1), l(4 * (size_t)g->n_walks + 1);
                const int r1 = phi_shard_walk_regions(g, 4, b.data());
                const int r2 = r1 ? r1 : phi_shard_slice_walks_all(g, 31, 25, 4, b.data(), f.data(), l.data());
                printf(" vtx=%u walks=%u regions rc=%d slices rc=%d", g->n_vtx, g->n_walks, r1, r2);
                phi_host_graph_free(G);
            }
            printf("\n");
        }
    }
    return 0;
}

struct Part { std::vector<uint32_t> rank_off, moff; std::vector<uint8_t> glen; std::vector<int32_t> gvtx; std::vector<uint16_t> mw; std::vector<uint64_t> ones; phi_index_result r; };

static int merge()
{
    std::mt19937_64 rng(5);
    const int NS = 120000, NW = 50, P = 5;
    struct Grp { int rank; std::vector<int32_t> v; std::vector<uint16_t> m; };
    std::vector<Grp> G;                                                   // the whole: groups of a rank in key order (first vertex 100 + j)
    for (int r = 0; r < NS; ++r) {
        const int ng = (rng() % 10 < 2) ? 1 + (int)(rng() % 3) : 0;
        for (int g = 0; g < ng; ++g) {
            Grp x; x.rank = r; x.v.push_back(100 + g);
            for (int i = (int)(rng() % 4); i > 0; --i) x.v.push_back(1000 + (int32_t)(rng() % 9000));
            for (int w = 0; w < NW; ++w) if (rng() % 3 == 0) x.m.push_back((uint16_t)w);
            if (x.m.empty()) x.m.push_back(7);
            G.push_back(x);
        }
    }
    std::vector<Part> parts(P);
    for (int p = 0; p < P; ++p) { parts[p].rank_off.assign(NS + 1, 0); parts[p].moff.push_back(0); parts[p].ones.assign(NW, 1); }
    for (size_t gi = 0; gi < G.size(); ++gi) {                           // two ranks out of three live on one part, the third is dealt out member by member
        std::vector<std::vector<uint16_t> > mem(P);
        const bool one_part = (G[gi].rank % 3) != 0;
        for (uint16_t w : G[gi].m) mem[one_part ? G[gi].rank % P : rng() % P].push_back(w);
        for (int p = 0; p < P; ++p) if (!mem[p].empty()) {
            Part &q = parts[p];
            q.rank_off[G[gi].rank + 1]++; q.glen.push_back((uint8_t)G[gi].v.size()); q.gvtx.insert(q.gvtx.end(), G[gi].v.begin(), G[gi].v.end());
            q.mw.insert(q.mw.end(), mem[p].begin(), mem[p].end()); q.moff.push_back((uint32_t)q.mw.size());
        }
    }
    std::vector<const phi_index_result *> ptr;
    for (int p = 0; p < P; ++p) {
        Part &q = parts[p];
        for (int r = 0; r < NS; ++r) q.rank_off[r + 1] += q.rank_off[r];
        memset(&q.r, 0, sizeof q.r);
        q.r.count_sp_r = NS; q.r.n_walks = NW; q.r.n_groups = q.glen.size(); q.r.n_anchors = q.mw.size(); q.r.n_group_vtx = q.gvtx.size();
        q.r.rank_off = q.rank_off.data(); q.r.group_len = q.glen.data(); q.r.group_vtx = q.gvtx.data(); q.r.group_member_off = q.moff.data();
        q.r.member_walk16 = q.mw.data(); q.r.minimizers_per_walk = q.ones.data(); q.r.anchors_per_walk = q.ones.data();
        ptr.push_back(&q.r);
    }
    int bad = 0;
    const char *threads[] = {"1", "3", "8"};
    for (int t = 0; t < 3; ++t) {
        setenv("PHI_MERGE_THREADS", threads[t], 1);
        phi_index_result *out = 0;
        const int rc = phi_index_result_merge(ptr.data(), P, &out);
        if (rc) { printf("merge rc %d\n", rc); return 1; }
        size_t gi = 0, vi = 0; bool ok = out->n_groups == G.size();
        for (int r = 0; r < NS && ok; ++r)
            for (uint32_t g = out->rank_off[r]; g < out->rank_off[r + 1] && ok; ++g, ++gi) {
                ok = gi < G.size() && G[gi].rank == r && out->group_len[g] == G[gi].v.size() && !memcmp(out->group_vtx + vi, G[gi].v.data(), 4 * G[gi].v.size());
                vi += out->group_len[g];
                const uint32_t m0 = out->group_member_off[g], m1 = out->group_member_off[g + 1];
                ok = ok && m1 - m0 == G[gi].m.size() && !memcmp(out->member_walk16 + m0, G[gi].m.data(), 2 * G[gi].m.size());
            }
        ok = ok && out->minimizers_per_walk[3] == (uint64_t)P;
        printf("merge on %s thread(s): %s (%zu groups)\n", threads[t], ok ? "identical to the whole" : "MISMATCH", G.size());
        bad += ok ? 0 : 1;
        phi_gpu_index_result_free(out);
    }
    return bad;
}

int main(int argc, char **argv)
{
    if (argc >= 2 && !strcmp(argv[1], "merge")) return merge();
    if (argc >= 3 && !strcmp(argv[1], "load")) return load(argc - 2, argv + 2);
    fprintf(stderr, "usage: %s load <file>... | merge\n", argv[0]);
    return 2;
}
