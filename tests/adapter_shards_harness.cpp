// TEST INFRASTRUCTURE (CPU): drives integration/phi_shards.hpp — the per-GPU views the reference-side adapter builds for a
// multi-GPU run — on a GFA + read file and dumps every shard, so that tests/test_adapter_shards.py can compare them with the
// library's own partition helpers rank by rank.      usage: harness graph.gfa reads.fa W k w out.bin
#include "phi_shards.hpp"

#include <cstdio>
#include <cstdlib>

static void put(FILE *f, const void *p, size_t n) { if (n && fwrite(p, 1, n, f) != n) { perror("write"); exit(2); } }
static void put64(FILE *f, uint64_t v) { put(f, &v, 8); }

int main(int argc, char **argv)
{
    if (argc != 7) { fprintf(stderr, "usage: %s graph.gfa reads W k w out.bin\n", argv[0]); return 2; }
    char err[512];
    phi_host_graph *G = 0; phi_host_reads *R = 0;
    if (phi_host_graph_load(argv[1], &G, err, sizeof err) != PHI_OK) { fprintf(stderr, "%s\n", err); return 1; }
    if (phi_host_reads_load(argv[2], &R, err, sizeof err) != PHI_OK) { fprintf(stderr, "%s\n", err); return 1; }
    const int W = atoi(argv[3]), k = atoi(argv[4]), w = atoi(argv[5]);
    std::vector<phi_adapter::detail::Shard> shards;
    const bool by_region = phi_adapter::detail::build_shards(*phi_host_graph_view(G), *phi_host_reads_view(R), W, k, w, shards);
    FILE *f = fopen(argv[6], "wb");
    if (!f) { perror(argv[6]); return 2; }
    put64(f, (uint64_t)W); put64(f, by_region ? 1 : 0);
    for (int r = 0; r < W; ++r) {
        const phi_adapter::detail::Shard &s = shards[r];
        const uint64_t steps = s.g.n_walks ? s.g.walk_off[s.g.n_walks] : 0, bases = s.rd.n_reads ? s.rd.read_off[s.rd.n_reads] : 0;
        put64(f, s.region ? 1 : 0); put64(f, s.coord_lo); put64(f, s.coord_hi); put64(f, s.walk_id_base);
        put64(f, s.g.n_vtx); put64(f, s.g.n_walks); put64(f, steps); put64(f, s.rd.n_reads); put64(f, bases);
        put(f, s.g.walk_off, ((size_t)s.g.n_walks + 1) * 8); put(f, s.g.walk_vtx, (size_t)steps * 4);
        put(f, s.rd.read_off, ((size_t)s.rd.n_reads + 1) * 8); put(f, s.rd.read_bases, (size_t)bases);
    }
    fclose(f);
    phi_host_graph_free(G); phi_host_reads_free(R);
    return 0;
}
