// TEST INFRASTRUCTURE ONLY — links against the UNMODIFIED reference objects.
//
// Calls the reference's own public entry points
//   gfa_read                      (/root/reference/src/gfa-io.cpp:462)
//   ILP_index::read_gfa           (/root/reference/src/ILP_index.cpp:20)
//   ILP_index::read_ip_reads      (/root/reference/src/ILP_index.cpp:313)
//   ILP_index::index_kmers(h)     (/root/reference/src/ILP_index.cpp:359)
//   ILP_index::compute_hashes(r)  (/root/reference/src/ILP_index.cpp:447)
// and writes (a) the flat graph/read views exactly as the reference numbers
// them and (b) the per-walk minimizer lists and per-read hash sets, as a
// sequence of named arrays that tests/phi_io.py reads.  Used to pin oracle/ and
// to generate tests/golden/.  Never part of the product.
//
// usage: ref_probe -g graph.gfa[.gz] -r reads.f[aq][.gz] -o out.phiarr [-k K] [-w W] [-t T] [--graph-only]
#include "gfa-priv.h"
#include "ILP_index.h"
#include <cstring>

static FILE *g_out;

static void put_array(const char *name, char dtype, const void *data, uint64_t n, size_t elt)
{
    uint32_t nl = (uint32_t)strlen(name);
    fwrite(&nl, 4, 1, g_out);
    fwrite(name, 1, nl, g_out);
    fwrite(&dtype, 1, 1, g_out);
    fwrite(&n, 8, 1, g_out);
    if (n) fwrite(data, elt, n, g_out);
}
template <class T> static void put(const char *name, char dtype, const std::vector<T> &v)
{
    put_array(name, dtype, v.data(), v.size(), sizeof(T));
}

int main(int argc, char **argv)
{
    std::string gfa_file, reads_file, out_file;
    int k = 31, w = 25, threads = 4, graph_only = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-g" && i + 1 < argc) gfa_file = argv[++i];
        else if (a == "-r" && i + 1 < argc) reads_file = argv[++i];
        else if (a == "-o" && i + 1 < argc) out_file = argv[++i];
        else if (a == "-k" && i + 1 < argc) k = atoi(argv[++i]);
        else if (a == "-w" && i + 1 < argc) w = atoi(argv[++i]);
        else if (a == "-t" && i + 1 < argc) threads = atoi(argv[++i]);
        else if (a == "--graph-only") graph_only = 1;
    }
    if (gfa_file.empty() || out_file.empty()) {
        fprintf(stderr, "usage: ref_probe -g graph.gfa -r reads -o out.phiarr [-k K] [-w W] [-t T] [--graph-only]\n");
        return 1;
    }
    mg_realtime0 = realtime();
    gfa_t *g = gfa_read(gfa_file.c_str());
    if (!g) { fprintf(stderr, "ref_probe: cannot read %s\n", gfa_file.c_str()); return 1; }
    ILP_index *ix = new ILP_index(g);
    ix->read_gfa();
    ix->k_mer = k; ix->window = w; ix->num_threads = threads;

    std::vector<std::pair<std::string, std::string> > reads;
    if (!reads_file.empty()) ix->read_ip_reads(reads, reads_file);

    g_out = fopen(out_file.c_str(), "wb");
    if (!g_out) { perror(out_file.c_str()); return 1; }

    std::vector<int32_t> params; params.push_back(k); params.push_back(w);
    put("params_kw", 'i', params);

    // flat graph view (vertex ids / walk ids as the reference numbers them)
    std::vector<uint64_t> seg_off(1, 0); std::vector<uint8_t> seg_bases;
    for (uint32_t v = 0; v < ix->n_vtx; ++v) {
        seg_bases.insert(seg_bases.end(), ix->node_seq[v].begin(), ix->node_seq[v].end());
        seg_off.push_back(seg_bases.size());
    }
    put("seg_off", 'Q', seg_off); put("seg_bases", 'B', seg_bases);
    std::vector<uint64_t> walk_off(1, 0); std::vector<uint32_t> walk_vtx;
    std::vector<uint8_t> names;
    for (uint32_t h = 0; h < ix->num_walks; ++h) {
        walk_vtx.insert(walk_vtx.end(), ix->paths[h].begin(), ix->paths[h].end());
        walk_off.push_back(walk_vtx.size());
        names.insert(names.end(), ix->hap_id2name[h].begin(), ix->hap_id2name[h].end());
        names.push_back('\n');
    }
    put("walk_off", 'Q', walk_off); put("walk_vtx", 'I', walk_vtx); put("walk_names", 'B', names);
    put("top_order_map", 'i', ix->top_order_map);
    std::vector<uint64_t> adj_off(1, 0); std::vector<uint32_t> adj;
    for (uint32_t v = 0; v < ix->n_vtx; ++v) {
        adj.insert(adj.end(), ix->adj_list[v].begin(), ix->adj_list[v].end());
        adj_off.push_back(adj.size());
    }
    put("adj_off", 'Q', adj_off); put("adj", 'I', adj);

    std::vector<uint64_t> read_off(1, 0); std::vector<uint8_t> read_bases;
    for (size_t r = 0; r < reads.size(); ++r) {
        read_bases.insert(read_bases.end(), reads[r].second.begin(), reads[r].second.end());
        read_off.push_back(read_bases.size());
    }
    put("read_off", 'Q', read_off); put("read_bases", 'B', read_bases);

    if (!graph_only) {
        // per-walk minimizers: reference index_kmers, called exactly as ILP_function does (:561)
        std::vector<std::vector<std::pair<uint64_t, Anchor> > > km(ix->num_walks);
        #pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
        for (int32_t h = 0; h < (int32_t)ix->num_walks; ++h) km[h] = ix->index_kmers(h);
        std::vector<uint64_t> wm_off(1, 0), wm_hash, wm_voff(1, 0); std::vector<int32_t> wm_vtx;
        for (uint32_t h = 0; h < ix->num_walks; ++h) {
            for (size_t i = 0; i < km[h].size(); ++i) {
                wm_hash.push_back(km[h][i].first);
                wm_vtx.insert(wm_vtx.end(), km[h][i].second.k_mers.begin(), km[h][i].second.k_mers.end());
                wm_voff.push_back(wm_vtx.size());
            }
            wm_off.push_back(wm_hash.size());
        }
        put("wm_off", 'Q', wm_off); put("wm_hash", 'Q', wm_hash); put("wm_voff", 'Q', wm_voff); put("wm_vtx", 'i', wm_vtx);

        // per-read hash sets: reference compute_hashes (:620); it upper-cases its argument in place
        std::vector<std::set<uint64_t> > rh(reads.size());
        #pragma omp parallel for num_threads(threads) schedule(dynamic, 64)
        for (int64_t r = 0; r < (int64_t)reads.size(); ++r) rh[r] = ix->compute_hashes(reads[r].second);
        std::vector<uint64_t> rh_off(1, 0), rh_hash;
        for (size_t r = 0; r < reads.size(); ++r) {
            rh_hash.insert(rh_hash.end(), rh[r].begin(), rh[r].end());
            rh_off.push_back(rh_hash.size());
        }
        put("rh_off", 'Q', rh_off); put("rh_hash", 'Q', rh_hash);
    }
    fclose(g_out);
    fprintf(stderr, "ref_probe: wrote %s (%u vertices, %u walks, %zu reads)\n", out_file.c_str(), ix->n_vtx, ix->num_walks, reads.size());
    return 0;
}
