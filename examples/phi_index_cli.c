/*
 * phi_index_cli — the ILP_index front end as a stand-alone command, written against the C ABI only (include/phi_gpu_index.h):
 * what /root/reference/src/main.cpp + ILP_index::ILP_function do up to line 743, with the same flags and the same log lines.
 *
 *   phi_index_cli -g graph.gfa[.gz] -r reads.fa|fq[.gz] [-k 31] [-w 25] [-T 1.0] [-d 0] [-o result.bin]
 *
 * -o writes the result in the "PHIRES3" layout integration/phi_adapter_testhook.hpp reads (header of seven u64, then the arrays
 * of phi_index_result in declaration order, each padded to 8 bytes).  Plain C99: the same calls work from cgo / JNI / ctypes.
 * Build: gcc -std=c99 -O2 -pthread -Iinclude examples/phi_index_cli.c -o phi_index_cli -Lphi_b200 -lphi_gpu_index -Wl,-rpath,$PWD/phi_b200
 */
#define _POSIX_C_SOURCE 200809L
#include "phi_gpu_index.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static void put(FILE *f, const void *p, size_t bytes)
{
    static const char zero[8] = {0};
    if (bytes) fwrite(p, 1, bytes, f);
    if (bytes & 7) fwrite(zero, 1, 8 - (bytes & 7), f);
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

struct ctx_job { phi_gpu_index_ctx *ctx; int rc; char err[512]; double t_done; };
static void *create_ctx(void *p)
{
    struct ctx_job *j = (struct ctx_job *)p;
    j->rc = phi_gpu_index_create(-1, &j->ctx);
    /* without a ctx the library keeps the message per thread: fetch it on the thread that made the call */
    if (j->rc != PHI_OK) { strncpy(j->err, phi_gpu_last_error(NULL), sizeof j->err - 1); j->err[sizeof j->err - 1] = 0; }
    j->t_done = now_s();
    return NULL;
}
struct reads_job { const char *path; phi_host_reads *hr; int rc; char err[512]; double t_done; };
static void *load_reads(void *p)
{
    struct reads_job *j = (struct reads_job *)p;
    j->rc = phi_host_reads_load(j->path, &j->hr, j->err, sizeof j->err);
    j->t_done = now_s();
    return NULL;
}

int main(int argc, char **argv)
{
    const char *gfa = NULL, *reads = NULL, *out = NULL;
    phi_index_params prm = {31, 25, 1.0f, 0};
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "-g")) gfa = argv[i + 1];
        else if (!strcmp(argv[i], "-r")) reads = argv[i + 1];
        else if (!strcmp(argv[i], "-o")) out = argv[i + 1];
        else if (!strcmp(argv[i], "-k")) prm.k = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-w")) prm.w = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-T")) prm.threshold = (float)atof(argv[i + 1]);
        else if (!strcmp(argv[i], "-d")) prm.debug = atoi(argv[i + 1]);
        else { fprintf(stderr, "unknown option %s\n", argv[i]); return 1; }
    }
    if (!gfa || !reads) { fprintf(stderr, "usage: phi_index_cli -g graph.gfa -r reads.fq [-k 31] [-w 25] [-T 1.0] [-d 0] [-o result.bin]\n"); return 1; }

    /* three things that do not depend on each other run side by side: the CUDA context (a few hundred ms in a fresh process), the
     * read file and — on this thread — the graph file */
    const double t0 = now_s();
    struct ctx_job cj; cj.ctx = NULL; cj.rc = PHI_OK; cj.err[0] = 0; cj.t_done = t0;
    struct reads_job rj; rj.path = reads; rj.hr = NULL; rj.rc = PHI_OK; rj.err[0] = 0; rj.t_done = t0;
    pthread_t t_ctx, t_reads;
    const int have_ctx_thread = pthread_create(&t_ctx, NULL, create_ctx, &cj) == 0;
    const int have_reads_thread = pthread_create(&t_reads, NULL, load_reads, &rj) == 0;
    char err[512];
    phi_host_graph *hg = NULL;
    const int grc = phi_host_graph_load(gfa, &hg, err, sizeof err);
    const double t_graph = now_s();
    if (have_reads_thread) pthread_join(t_reads, NULL); else load_reads(&rj);
    if (grc != PHI_OK || rj.rc != PHI_OK) {
        fprintf(stderr, "Error: %s\n", grc != PHI_OK ? err : rj.err);
        if (have_ctx_thread) pthread_join(t_ctx, NULL);
        return 1;
    }
    phi_host_reads *hr = rj.hr;
    const phi_graph_view *g = phi_host_graph_view(hg);
    const phi_reads_view *rd = phi_host_reads_view(hr);
    fprintf(stderr, "Graph has %u vertices, %u walks and read has %llu reads\n", g->n_vtx, g->n_walks, (unsigned long long)rd->n_reads);

    if (have_ctx_thread) pthread_join(t_ctx, NULL); else create_ctx(&cj);
    phi_gpu_index_ctx *ctx = cj.ctx; phi_index_result *res = NULL;
    int rc = cj.rc;
    if (rc != PHI_OK) { fprintf(stderr, "Error: GPU ILP_index front end failed (%d): %s\n", rc, cj.err); return 1; }
    const double t_ready = now_s();
    rc = phi_gpu_index_run(ctx, g, rd, &prm, &res);
    const double t_run = now_s();
    if (rc != PHI_OK) { fprintf(stderr, "Error: GPU ILP_index front end failed (%d): %s\n", rc, phi_gpu_last_error(ctx)); return 1; }

    fprintf(stderr, "Number of Minimizers\n");
    for (uint32_t h = 0; h < g->n_walks; ++h) fprintf(stderr, "%s : %d\n", phi_host_graph_walk_name(hg, h), (int)res->minimizers_per_walk[h]);
    fprintf(stderr, "Indexed reads with spectrum size: %d\n", res->count_sp_r);
    fprintf(stderr, "Number of Anchors\n");
    for (uint32_t h = 0; h < g->n_walks; ++h) fprintf(stderr, "%s : %d\n", phi_host_graph_walk_name(hg, h), (int)res->anchors_per_walk[h]);
    const long long filtered = res->n_filtered, retained = res->count_sp_r - filtered;
    fprintf(stderr, "Filtered/Retained Minimizers: %.2f/%.2f%%\n", (float)filtered / (float)res->count_sp_r * 100, (float)retained / (float)res->count_sp_r * 100);
    phi_stage_times t;
    if (phi_gpu_index_last_times(ctx, &t) == PHI_OK)
        fprintf(stderr, "GPU front end: %.3f ms (copies in %.3f ms, out %.3f ms; %llu kernel launches); %llu groups, %llu anchors\n", t.total_ms, t.h2d_ms,
                t.d2h_ms, (unsigned long long)t.kernel_launches, (unsigned long long)res->n_groups, (unsigned long long)res->n_anchors);

    if (getenv("PHI_CLI_TIMES"))                     /* wall clock from program start: what ran side by side, and what the result waited for */
        fprintf(stderr, "[phi_index_cli] graph loaded %.3f s, reads loaded %.3f s, CUDA context ready %.3f s | all three ready %.3f s, front end done %.3f s "
                        "(phi_gpu_index_run %.3f s)\n", t_graph - t0, rj.t_done - t0, cj.t_done - t0, t_ready - t0, t_run - t0, t_run - t_ready);

    if (out) {
        FILE *f = fopen(out, "wb");
        if (!f) { fprintf(stderr, "Error: cannot write %s\n", out); return 1; }
        const uint64_t mb = res->member_walk16 ? 2 : 4;
        const uint64_t head[7] = {(uint64_t)res->count_sp_r, res->n_walks, (uint64_t)res->n_filtered, res->n_anchors, res->n_groups, res->n_group_vtx, mb};
        fwrite("PHIRES3\0", 1, 8, f); fwrite(head, 8, 7, f);
        put(f, res->spectrum, 8 * (size_t)res->count_sp_r);
        put(f, res->rank_off, 4 * ((size_t)res->count_sp_r + 1));
        put(f, res->group_len, res->n_groups);
        put(f, res->group_vtx, 4 * res->n_group_vtx);
        put(f, res->group_member_off, 4 * (res->n_groups + 1));
        if (mb == 2) put(f, res->member_walk16, 2 * res->n_anchors); else put(f, res->member_walk32, 4 * res->n_anchors);
        put(f, res->minimizers_per_walk, 8 * (size_t)res->n_walks);
        put(f, res->anchors_per_walk, 8 * (size_t)res->n_walks);
        fclose(f);
    }
    phi_gpu_index_result_free(res);
    phi_gpu_index_destroy(ctx);
    phi_host_reads_free(hr);
    phi_host_graph_free(hg);
    return 0;
}
