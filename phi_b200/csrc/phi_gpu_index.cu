// Host side of libphi_gpu_index.so: the C ABI declared in include/phi_gpu_index.h and the stage
// pipeline that replaces /root/reference/src/ILP_index.cpp:543-743.  C++ only (no PyTorch); one ctx
// drives one GPU: the calling thread owns the main stream (reads, spectrum, walk sketch, filter, result), a second
// thread it starts per run owns the second stream (graph preparation), a copy stream carries the uploads and the early
// part of the download.  There is no CPU fallback: without a usable device phi_gpu_index_create() fails and nothing
// else can be called.
#include "../../include/phi_gpu_index.h"
#include "kernels.h"
#include "result_box.h"

#include <algorithm>
#include <dlfcn.h>
#include <nccl.h>          // types and prototypes only: the functions are resolved with dlopen at comm_init
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

using namespace phi;

namespace {

thread_local std::string g_create_error = "no error";

struct DevBuf {                       // grow-only device buffer
    void *p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};


enum { EV_START, EV_H2D, EV_PREP0, EV_PREP, EV_RD0, EV_READS, EV_SPECTRUM, EV_WALKS, EV_FILTER, EV_END, EV_RK0, EV_RK1, EV_WK0, EV_WK1, EV_XS0, EV_XS1, EV_XH0, EV_XH1, EV_XH2, EV_COUNT };

}  // namespace

struct phi_gpu_index_ctx {
    int device = 0;
    cudaStream_t st = nullptr;             // main stream: reads, spectrum, walk sketch, filter, result copies
    cudaStream_t st2 = nullptr;            // second stream: graph upload + graph preparation / chunking, concurrent with the read stage
    cudaEvent_t ev[EV_COUNT] = {};
    cudaEvent_t ev_sync = nullptr;         // cross-stream ordering (no timing)
    // host -> device copies of phi_gpu_index_run: one copy stream, graph first (the second stream starts preparing it), then the
    // reads in pieces (the main stream sketches piece p while piece p+1 is on the wire)
    cudaStream_t st_copy = nullptr;
    cudaEvent_t ev_graph_in = nullptr, ev_piece[8] = {}, ev_wpiece[8] = {};
    int n_pieces = 0;                      // > 0: an upload is in flight and the read stage has to wait piece by piece
    uint64_t piece_end[8] = {};            // read bases [0, piece_end[p]) are on the device once ev_piece[p] has fired
    int n_wpieces = 0;                     // the walk steps arrive in this many pieces (after everything else of the graph: ev_graph_in)
    uint64_t wpiece_end[8] = {};           // walk steps [0, wpiece_end[p]) are on the device once ev_wpiece[p] has fired
    std::string err = "no error";
    uint64_t launches = 0;
    phi_stage_times times = {};

    // resident inputs
    bool have_inputs = false;
    uint32_t n_vtx = 0, n_walks = 0; uint64_t n_steps = 0, seg_total = 0, n_reads = 0, read_total = 0;
    DevBuf seg_off, seg_bases, walk_off, walk_vtx, top_order, read_off, read_bases;
    std::vector<uint64_t> h_walk_off;

    // work buffers
    DevBuf step_len, gbase, step_base, walk_len, tile_first_read, scan_scr, ctr;
    DevBuf walk_off_c, walk_vtx_c, flags64;
    DevBuf table, tblk, spec_a, spec_b, sort_scr, dir;
    uint32_t *h_tot = nullptr;             // pinned: occupied slots of the spectrum table
    DevBuf mpw, hit_rank, hit_chunk, hit_pos, hit_voff, hit_nv, hit_hash, vtx_pool;
    // walk chunks (chunks.cu): boundaries, fingerprints, representatives, tiles, hit segments, expanded survivors
    DevBuf tlen, tprefix, coord, cflags, cpos, chunk_step, c_walk, c_L, c_R, c_lo, c_hi, c_h1, c_h2, c_slot, c_rep, c_ninst, c_ntile, c_tile_base, ctable, tiles;
    DevBuf hseg_off, hseg_cnt, c_emitted, c_hits, c_surv, c_surv_vtx, member_cnt, member_off, x_rank, x_walk, x_pos, x_voff, x_nv, x_hash;
    uint32_t n_chunks = 0, n_tiles = 0; uint64_t unique_windows = 0, active_chunks = 0, rep_chunks = 0, unique_hits = 0, path_pos = 0; int dedupe = 1, chunk_shift = 11;
    uint64_t own_lo = 0, own_hi = ~0ull;   // owned range of the topological base coordinate (phi_gpu_index_set_walk_region); default: everything
    DevBuf g_slot, probe, rank_drop, flags, keys_a, keys_b, vals_a, vals_b, big_list, tmp_order, nv_out;
    DevBuf anchor_off, rank_off, anchor_len, anchor_walk, anchor_vtx, apw, dbg_hist;
    // grouped result (groups.cu): member walks of the representative chunks, slot / sub-offset of every hit, group sizes and offsets
    DevBuf fs_state;                       // ticket + one look-back word per tile of the fused step kernel
    DevBuf cm_off, cm_cursor, cm_tmp, cm_walk, hit_slot, hit_sub, hit_slot2, g_slot2, probe2, grp_cnt, grp_moff, grp_voff, members_tmp;
    unsigned long long *h_ctr = nullptr;   // pinned mirror of the counter block
    // what the graph-preparation thread (second stream) uses instead of ctr / h_ctr / scan_scr / flags / flags64
    DevBuf ctr2, scan_scr2, flags2, flags64_2, nv_out2; unsigned long long *h_ctr2 = nullptr;
    std::vector<PinnedBuf> pinned_pool;    // free pinned buffers (returned by phi_gpu_index_result_free)

    // multi-GPU (set by comm_init)
    int rank = 0, world = 1; uint32_t walk_id_base = 0, n_walks_global = 0;
    void *comm = nullptr;
    uint64_t gcap_hint = 0, gcap_hint2 = 0;   // group-table sizes that worked last time (local table, owner-side table)
    uint64_t spec_hint = 0, spec_hint_bases = 0;   // distinct read minimizers of the last run and the read bases they came from (spectrum table sizing)
    DevBuf xk_a, xk_b, xcnt, xoff, ag_send, ag_recv, m_rank, m_cnt, m_voff, m_nv, r_rank, r_walk, r_pos, r_voff, r_nv, r_vtx, s_rank, s_walk, s_pos, s_voff, s_nv, s_vtx;
    std::vector<uint64_t> own_off;       // [world + 1] first global rank owned by each GPU (multi-GPU runs)
    uint64_t *h_words = nullptr; size_t h_words_cap = 0;   // pinned: gathered words of the small collectives
    uint64_t *h_route = nullptr;         // pinned [512]: small host -> device parameter blocks of the record exchange

    size_t l2_persist_bytes = 0, l2_window_max = 0;   // persisting L2 set aside at create (0: not available)
    std::string err2;                      // error text of the graph-preparation thread (moved into err when its failure is reported)
    uint64_t launches2 = 0;                // kernels launched by that thread
    int fail(int code, const std::string &m) { err = m; return code; }
    int fail2(int code, const std::string &m) { err2 = m; return code; }
};

static std::mutex g_live_mu;                       // live ctxs: a result freed after its ctx releases its pinned buffers itself
static std::set<phi_gpu_index_ctx *> g_live_ctx;

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx->fail(PHI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
// the same for the graph-preparation thread (its own error text: the main thread may be failing at the same time)
#define CUP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx->fail2(PHI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

extern "C" int phi_gpu_index_abi_version(void) { return PHI_GPU_INDEX_ABI_VERSION; }

extern "C" const char *phi_gpu_last_error(const phi_gpu_index_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int phi_gpu_index_create(int device, phi_gpu_index_ctx **out)
{
    if (!out) { g_create_error = "out is NULL"; return PHI_ERR_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (this library has no CPU fallback)";
        return PHI_ERR_CUDA;
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= n) { g_create_error = "device index out of range"; return PHI_ERR_ARG; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return PHI_ERR_CUDA; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return PHI_ERR_CUDA; }
    if (prop.major != 10) {
        g_create_error = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + "; this library is built for sm_100a only";
        return PHI_ERR_CUDA;
    }
    phi_gpu_index_ctx *ctx = new phi_gpu_index_ctx();
    ctx->device = device;
    // Opt-in (PHI_GPU_L2_PIN=1): an L2 set-aside with persisting access windows over the vertex records (step kernel) and the radix
    // directory (walk kernel).  Measured on B200 (profiles/r2_variants_ab.txt): 48.94 vs 49.19 ms on c4, 2.13 vs 2.08 ms on c2 — the
    // set-aside takes L2 away from everything else, so it stays off.
    if (getenv("PHI_GPU_L2_PIN") && atoi(getenv("PHI_GPU_L2_PIN")) && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
        const size_t want = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, (size_t)prop.l2CacheSize / 2);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) { ctx->l2_persist_bytes = want; ctx->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize; }
        else cudaGetLastError();
    }
    if (const char *e_shift = getenv("PHI_GPU_CHUNK_SHIFT")) { int v = atoi(e_shift); if (v >= 4 && v <= 24) ctx->chunk_shift = v; }   // tuning only: results never depend on it
    if ((e = cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return PHI_ERR_CUDA; }
    // (a higher priority for this stream was tried: the preparation then finishes earlier but the read kernel, which fills every SM,
    // is slowed by exactly as much: the two together take the sum of their times either way)
    if ((e = cudaStreamCreateWithFlags(&ctx->st2, cudaStreamNonBlocking)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return PHI_ERR_CUDA; }
    for (int i = 0; i < EV_COUNT; ++i) cudaEventCreate(&ctx->ev[i]);
    cudaEventCreateWithFlags(&ctx->ev_sync, cudaEventDisableTiming);
    if ((e = cudaStreamCreateWithFlags(&ctx->st_copy, cudaStreamNonBlocking)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return PHI_ERR_CUDA; }
    cudaEventCreateWithFlags(&ctx->ev_graph_in, cudaEventDisableTiming);
    for (int i = 0; i < 8; ++i) { cudaEventCreateWithFlags(&ctx->ev_piece[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_wpiece[i], cudaEventDisableTiming); }
    if ((e = cudaHostAlloc((void **)&ctx->h_ctr, CTR_COUNT * 8, cudaHostAllocDefault)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return PHI_ERR_CUDA; }
    if ((e = cudaHostAlloc((void **)&ctx->h_tot, 64, cudaHostAllocDefault)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return PHI_ERR_CUDA; }
    if ((e = cudaHostAlloc((void **)&ctx->h_ctr2, CTR_COUNT * 8, cudaHostAllocDefault)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return PHI_ERR_CUDA; }
    if ((e = ctx->ctr.reserve(CTR_COUNT * 8)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return PHI_ERR_CUDA; }
    if ((e = ctx->ctr2.reserve(CTR_COUNT * 8)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return PHI_ERR_CUDA; }
    if ((e = cudaHostAlloc((void **)&ctx->h_route, 512 * 8, cudaHostAllocDefault)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return PHI_ERR_CUDA; }
    { std::lock_guard<std::mutex> lk(g_live_mu); g_live_ctx.insert(ctx); }
    *out = ctx;
    return PHI_OK;
}

static void comm_release(phi_gpu_index_ctx *ctx, bool abort);

extern "C" void phi_gpu_index_destroy(phi_gpu_index_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    if (ctx->st2) cudaStreamSynchronize(ctx->st2);
    comm_release(ctx, false);
    DevBuf *bufs[] = {&ctx->ctr2, &ctx->scan_scr2, &ctx->flags2, &ctx->flags64_2, &ctx->nv_out2,
                      &ctx->seg_off, &ctx->seg_bases, &ctx->walk_off, &ctx->walk_vtx, &ctx->top_order, &ctx->read_off, &ctx->read_bases,
                      &ctx->step_len, &ctx->gbase, &ctx->step_base, &ctx->walk_len,
                      &ctx->tile_first_read, &ctx->scan_scr, &ctx->ctr, &ctx->walk_off_c, &ctx->walk_vtx_c, &ctx->flags64, &ctx->table, &ctx->tblk,
                      &ctx->spec_a, &ctx->spec_b, &ctx->sort_scr, &ctx->dir, &ctx->mpw, &ctx->hit_rank, &ctx->hit_chunk, &ctx->hit_pos,
                      &ctx->tlen, &ctx->tprefix, &ctx->coord, &ctx->cflags, &ctx->cpos, &ctx->chunk_step, &ctx->c_walk, &ctx->c_L, &ctx->c_R, &ctx->c_lo,
                      &ctx->c_hi, &ctx->c_h1, &ctx->c_h2, &ctx->c_slot, &ctx->c_rep, &ctx->c_ninst, &ctx->c_ntile, &ctx->c_tile_base, &ctx->ctable,
                      &ctx->tiles, &ctx->hseg_off, &ctx->hseg_cnt, &ctx->c_emitted, &ctx->c_hits, &ctx->c_surv, &ctx->c_surv_vtx, &ctx->member_cnt, &ctx->member_off,
                      &ctx->x_rank, &ctx->x_walk, &ctx->x_pos, &ctx->x_voff, &ctx->x_nv, &ctx->x_hash,
                      &ctx->xk_a, &ctx->xk_b, &ctx->xcnt, &ctx->xoff, &ctx->ag_send, &ctx->ag_recv, &ctx->m_rank, &ctx->m_cnt, &ctx->m_voff, &ctx->m_nv,
                      &ctx->r_rank, &ctx->r_walk, &ctx->r_pos, &ctx->r_voff, &ctx->r_nv, &ctx->r_vtx, &ctx->s_rank, &ctx->s_walk, &ctx->s_pos,
                      &ctx->s_voff, &ctx->s_nv, &ctx->s_vtx,
                      &ctx->hit_voff, &ctx->hit_nv, &ctx->hit_hash, &ctx->vtx_pool, &ctx->g_slot, &ctx->probe, &ctx->rank_drop, &ctx->flags,
                      &ctx->keys_a, &ctx->keys_b, &ctx->vals_a, &ctx->vals_b, &ctx->big_list, &ctx->tmp_order, &ctx->nv_out,
                      &ctx->fs_state, &ctx->cm_off, &ctx->cm_cursor, &ctx->cm_tmp, &ctx->cm_walk, &ctx->hit_slot, &ctx->hit_sub, &ctx->hit_slot2, &ctx->g_slot2, &ctx->probe2,
                      &ctx->grp_cnt, &ctx->grp_moff, &ctx->grp_voff, &ctx->members_tmp,
                      &ctx->dbg_hist, &ctx->anchor_off, &ctx->rank_off, &ctx->anchor_len, &ctx->anchor_walk, &ctx->anchor_vtx, &ctx->apw};
    for (DevBuf *b : bufs) b->release();
    {
        std::lock_guard<std::mutex> lk(g_live_mu);
        g_live_ctx.erase(ctx);
        for (PinnedBuf &pb : ctx->pinned_pool) cudaFreeHost(pb.p);
        ctx->pinned_pool.clear();
    }
    if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
    if (ctx->h_ctr2) cudaFreeHost(ctx->h_ctr2);
    if (ctx->h_tot) cudaFreeHost(ctx->h_tot);
    if (ctx->h_words) cudaFreeHost(ctx->h_words);
    if (ctx->h_route) cudaFreeHost(ctx->h_route);
    for (int i = 0; i < EV_COUNT; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->ev_sync) cudaEventDestroy(ctx->ev_sync);
    if (ctx->ev_graph_in) cudaEventDestroy(ctx->ev_graph_in);
    for (int i = 0; i < 8; ++i) { if (ctx->ev_piece[i]) cudaEventDestroy(ctx->ev_piece[i]); if (ctx->ev_wpiece[i]) cudaEventDestroy(ctx->ev_wpiece[i]); }
    if (ctx->st_copy) { cudaStreamSynchronize(ctx->st_copy); cudaStreamDestroy(ctx->st_copy); }
    if (ctx->st) cudaStreamDestroy(ctx->st);
    if (ctx->st2) cudaStreamDestroy(ctx->st2);
    delete ctx;
}

static int check_views(phi_gpu_index_ctx *ctx, const phi_graph_view *g, const phi_reads_view *r)
{
    if (!g || !r) return ctx->fail(PHI_ERR_ARG, "graph/reads view is NULL");
    if ((g->n_vtx && (!g->seg_off || !g->top_order_map)) || (g->n_walks && !g->walk_off)) return ctx->fail(PHI_ERR_ARG, "graph view has NULL arrays");
    if (r->n_reads && !r->read_off) return ctx->fail(PHI_ERR_ARG, "reads view has NULL arrays");
    // offsets: start at 0, never decrease; data arrays present when the totals say so (O(n_vtx + n_walks + n_reads); the vertex ids of the
    // walk steps are range-checked on the device by the step pass, before anything indexes with them)
    auto monotone = [](const uint64_t *off, uint64_t n) { if (off[0] != 0) return false; for (uint64_t i = 0; i < n; ++i) if (off[i + 1] < off[i]) return false; return true; };
    if (g->n_vtx && !monotone(g->seg_off, g->n_vtx)) return ctx->fail(PHI_ERR_ARG, "graph view: seg_off must start at 0 and be non-decreasing");
    if (g->n_walks && !monotone(g->walk_off, g->n_walks)) return ctx->fail(PHI_ERR_ARG, "graph view: walk_off must start at 0 and be non-decreasing");
    if (r->n_reads && !monotone(r->read_off, r->n_reads)) return ctx->fail(PHI_ERR_ARG, "reads view: read_off must start at 0 and be non-decreasing");
    if (g->n_vtx && g->seg_off[g->n_vtx] && !g->seg_bases) return ctx->fail(PHI_ERR_ARG, "graph view: seg_bases is NULL");
    if (g->n_walks && g->walk_off[g->n_walks] && (!g->walk_vtx || !g->n_vtx)) return ctx->fail(PHI_ERR_ARG, "graph view: walk_vtx is NULL (or there are no vertices)");
    if (r->n_reads && r->read_off[r->n_reads] && !r->read_bases) return ctx->fail(PHI_ERR_ARG, "reads view: read_bases is NULL");
    return PHI_OK;
}

// Host -> device copies, issued without waiting: the reads on the main stream (the read stage starts right behind them),
// the graph on the second stream (the graph preparation follows it there).  The caller's buffers must stay valid until both
// streams have been synchronised.
static int upload_async(phi_gpu_index_ctx *ctx, const phi_graph_view *g, const phi_reads_view *r)
{
    int rc = check_views(ctx, g, r);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    ctx->have_inputs = false;
    ctx->n_vtx = g->n_vtx; ctx->n_walks = g->n_walks;
    ctx->seg_total = g->n_vtx ? g->seg_off[g->n_vtx] : 0;
    ctx->n_steps = g->n_walks ? g->walk_off[g->n_walks] : 0;
    ctx->n_reads = r->n_reads;
    ctx->read_total = r->n_reads ? r->read_off[r->n_reads] : 0;
    ctx->h_walk_off.assign(g->n_walks + 1, 0);
    if (g->n_walks) memcpy(ctx->h_walk_off.data(), g->walk_off, (size_t)(g->n_walks + 1) * 8);
    static const uint64_t zero_off[1] = {0};

    CU(ctx->seg_off.reserve(((size_t)g->n_vtx + 1) * 8));
    CU(ctx->seg_bases.reserve(ctx->seg_total + 64));
    CU(ctx->top_order.reserve((size_t)g->n_vtx * 4 + 4));
    CU(ctx->walk_off.reserve(((size_t)g->n_walks + 1) * 8));
    CU(ctx->walk_vtx.reserve(ctx->n_steps * 4 + 4));
    CU(ctx->read_off.reserve((ctx->n_reads + 1) * 8));
    CU(ctx->read_bases.reserve(ctx->read_total + 64));
    // sequence buffers carry 16 readable bytes in front and zero padding behind: the sketch kernels use unaligned 8-byte loads.
    // One copy stream, in the order the pipeline wants the data: graph first, then the reads in pieces.
    cudaStream_t sc = ctx->st_copy;
    CU(cudaMemcpyAsync(ctx->seg_off.p, g->n_vtx ? g->seg_off : zero_off, ((size_t)g->n_vtx + 1) * 8, cudaMemcpyHostToDevice, sc));
    if (g->n_vtx) CU(cudaMemcpyAsync(ctx->top_order.p, g->top_order_map, (size_t)g->n_vtx * 4, cudaMemcpyHostToDevice, sc));
    CU(cudaMemcpyAsync(ctx->walk_off.p, g->n_walks ? g->walk_off : zero_off, ((size_t)g->n_walks + 1) * 8, cudaMemcpyHostToDevice, sc));
    CU(cudaMemsetAsync(ctx->seg_bases.p, 0, 16, sc));
    CU(cudaMemsetAsync((char *)ctx->seg_bases.p + 16 + ctx->seg_total, 0, 32, sc));
    if (ctx->seg_total) CU(cudaMemcpyAsync((char *)ctx->seg_bases.p + 16, g->seg_bases, ctx->seg_total, cudaMemcpyHostToDevice, sc));
    CU(cudaEventRecord(ctx->ev_graph_in, sc));                               // everything of the graph but the walk steps
    {   // the walk steps (the bulk of a big graph) in pieces: the step pass of the graph preparation follows piece by piece
        const uint64_t S = ctx->n_steps, tile = walk_steps_fused_tile_steps();
        const int WP = S >= (8u << 20) ? 8 : 1;
        ctx->n_wpieces = WP;
        uint64_t lo = 0;
        for (int p = 0; p < WP; ++p) {
            uint64_t hi = p + 1 == WP ? S : (S * (p + 1) / WP) / tile * tile;
            if (hi > lo) CU(cudaMemcpyAsync(ctx->walk_vtx.as<uint32_t>() + lo, g->walk_vtx + lo, (hi - lo) * 4, cudaMemcpyHostToDevice, sc));
            ctx->wpiece_end[p] = hi; lo = hi;
            CU(cudaEventRecord(ctx->ev_wpiece[p], sc));
        }
    }
    CU(cudaMemcpyAsync(ctx->read_off.p, ctx->n_reads ? r->read_off : zero_off, (ctx->n_reads + 1) * 8, cudaMemcpyHostToDevice, sc));
    CU(cudaMemsetAsync(ctx->read_bases.p, 0, 16, sc));
    CU(cudaMemsetAsync((char *)ctx->read_bases.p + 16 + ctx->read_total, 0, 32, sc));
    const int P = ctx->read_total >= (8u << 20) ? 8 : 1;
    ctx->n_pieces = P;
    for (int p = 0; p < P; ++p) {
        const uint64_t lo = ctx->read_total * p / P, hi = ctx->read_total * (p + 1) / P;
        if (hi > lo) CU(cudaMemcpyAsync((char *)ctx->read_bases.p + 16 + lo, r->read_bases + lo, hi - lo, cudaMemcpyHostToDevice, sc));
        ctx->piece_end[p] = hi;
        CU(cudaEventRecord(ctx->ev_piece[p], sc));
    }
    ctx->have_inputs = true;
    return PHI_OK;
}

extern "C" int phi_gpu_index_upload(phi_gpu_index_ctx *ctx, const phi_graph_view *g, const phi_reads_view *r)
{
    if (!ctx) return PHI_ERR_ARG;
    int rc = upload_async(ctx, g, r);
    if (rc) { ctx->have_inputs = false; return rc; }
    CU(cudaStreamSynchronize(ctx->st_copy));
    ctx->n_pieces = 0;                                                      // resident: nothing to wait for
    return PHI_OK;
}

static int bits_for(uint64_t max_value)     // number of bits needed to represent values in [0, max_value]
{
    int b = 0;
    while (b < 64 && (max_value >> b)) ++b;
    return b ? b : 1;
}

static cudaError_t read_counters_on(phi_gpu_index_ctx *ctx, cudaStream_t st, DevBuf &ctr, unsigned long long *h_ctr)
{
    (void)ctx;
    cudaError_t e = cudaMemcpyAsync(h_ctr, ctr.p, CTR_COUNT * 8, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}

static cudaError_t read_counters(phi_gpu_index_ctx *ctx)
{
    cudaError_t e = cudaMemcpyAsync(ctx->h_ctr, ctx->ctr.p, CTR_COUNT * 8, cudaMemcpyDeviceToHost, ctx->st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(ctx->st);
}

__global__ void count_positions_kernel(const uint64_t *off, uint64_t n, int k, int w, unsigned long long *out)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    if (i < n) { uint64_t len = off[i + 1] - off[i]; if (len >= (uint64_t)(w + k - 1)) v = len - k + 1; }
    for (int d = 16; d; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

__global__ void count_survivors_kernel(const uint32_t *flags, uint64_t n, unsigned long long *ctr)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    bool s = i < n && flags[i];
    uint32_t b = __ballot_sync(0xFFFFFFFFu, s);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&ctr[CTR_SURVIVORS], (unsigned long long)__popc(b));
}

namespace {

struct RunOut {                    // device-side products of one run
    uint32_t n_spec = 0;
    int member_walk_bytes = 4;
    uint64_t n_hits = 0, n_hit_vtx = 0, n_surv = 0, n_groups = 0, n_anchor_vtx = 0;   // n_hits: hits of the representative chunks
    uint64_t path_hits = 0;                                             // hits of all walks
    uint64_t read_pos = 0, path_pos = 0, read_emitted = 0, path_emitted = 0;
    int64_t n_filtered = 0;
};

}  // namespace

// ---- stage: graph preparation: step base offsets, walk lengths, then the chunk table and the tiles of the representative
// chunks (depends on k and w through the chunk context)
static ChunkTable chunk_table(phi_gpu_index_ctx *ctx)
{
    ChunkTable C;
    C.n_chunks = ctx->n_chunks; C.chunk_step = ctx->chunk_step.as<uint32_t>();
    C.c_walk = ctx->c_walk.as<uint32_t>(); C.c_L = ctx->c_L.as<uint32_t>(); C.c_R = ctx->c_R.as<uint32_t>();
    C.c_lo = ctx->c_lo.as<uint32_t>(); C.c_hi = ctx->c_hi.as<uint32_t>(); C.c_h1 = ctx->c_h1.as<uint64_t>(); C.c_h2 = ctx->c_h2.as<uint64_t>();
    C.c_slot = ctx->c_slot.as<uint32_t>(); C.c_rep = ctx->c_rep.as<uint32_t>(); C.c_ninst = ctx->c_ninst.as<uint32_t>();
    C.c_ntile = ctx->c_ntile.as<uint32_t>(); C.c_tile_base = ctx->c_tile_base.as<uint32_t>();
    return C;
}

// Steps: packed lengths + chunk flags (one pass), one u64 scan, step bases / walk lengths / chunk table, fingerprints,
// grouping, tiles.  Two host-side waits (number of chunks; number of tiles + walk lengths + flags), on the second stream.
static int stage_graph_prep(phi_gpu_index_ctx *ctx, int k, int w, std::vector<uint64_t> &h_walk_len, const uint32_t *&d_walk_vtx,
                            const uint64_t *&d_walk_off, uint64_t &n_steps_eff, int &walks_monotone)
{
    // Runs on its own host thread (run_pipeline) next to the read stage and the spectrum exchange: it touches the second stream, that
    // stream's counter block and scratch buffers and the graph-side buffers only, counts its launches apart and keeps its own error text.
    cudaStream_t st = ctx->st2;
    DevBuf &ctr = ctx->ctr2, &scan_scr = ctx->scan_scr2, &flags = ctx->flags2, &flags64 = ctx->flags64_2;
    unsigned long long *h_ctr = ctx->h_ctr2;
    uint64_t *launches = &ctx->launches2;
    walks_monotone = 1;
    const uint32_t H = ctx->n_walks, V = ctx->n_vtx; uint64_t S = ctx->n_steps;
    d_walk_vtx = ctx->walk_vtx.as<uint32_t>(); d_walk_off = ctx->walk_off.as<uint64_t>(); n_steps_eff = S;
    h_walk_len.assign(H, 0);
    ctx->n_chunks = ctx->n_tiles = 0; ctx->unique_windows = ctx->active_chunks = ctx->rep_chunks = ctx->path_pos = 0;
    unsigned long long *d_ctr = ctr.as<unsigned long long>();
    if (ctx->n_pieces) CUP(cudaStreamWaitEvent(st, ctx->ev_graph_in, 0));   // phi_gpu_index_run: the graph is still on the wire
    CUP(cudaEventRecord(ctx->ev[EV_PREP0], st));
    CUP(cudaMemsetAsync(ctr.p, 0, CTR_COUNT * 8, st));
    memset(h_ctr, 0, CTR_COUNT * 8);
    if (!H || !S) return PHI_OK;
    if (S >= 0xFFFFFFFFull) return ctx->fail2(PHI_ERR_UNSUPPORTED, "more than 2^32-2 walk steps on one GPU; shard the walks over more GPUs");
    // topological base coordinate of every vertex (chunk boundaries are defined on it)
    CUP(ctx->tlen.reserve((size_t)V * 4 + 4)); CUP(ctx->tprefix.reserve(((size_t)V + 1) * 8)); CUP(ctx->coord.reserve((size_t)V * 16 + 16));
    CUP(scan_scr.reserve(std::max({scan_u32_to_u64_scratch((uint64_t)V + 1), scan_u32_to_u64_scratch(S + 1), (size_t)1024})));
    CUP(chunk_topo_coord(ctx->top_order.as<int32_t>(), ctx->seg_off.as<uint64_t>(), V, ctx->chunk_shift, ctx->own_lo, ctx->own_hi, ctx->tlen.as<uint32_t>(),
                        ctx->tprefix.as<uint64_t>(), ctx->coord.as<uint4>(), scan_scr.p, d_ctr, st, launches));
    // common case: one kernel for step lengths, chunk flags, their scan, step bases, walk lengths and the chunk starts
    const uint64_t S0 = S;
    CUP(ctx->step_base.reserve(S * 4 + 4)); CUP(ctx->walk_len.reserve((size_t)H * 8 + 8));
    CUP(ctx->chunk_step.reserve((S + 2) * 4)); CUP(ctx->c_walk.reserve((S + 2) * 4));       // at most one chunk per step
    CUP(ctx->fs_state.reserve(walk_steps_fused_tiles(S) * 8 + 16));
    CUP(cudaMemsetAsync(d_ctr + CTR_ZERO_STEPS, 0, 8, st)); CUP(cudaMemsetAsync(d_ctr + CTR_CHUNK_FLAGS, 0, 8, st));
    // the step kernel gathers one 16-byte vertex record per step while it streams the steps through L2: optionally keep the records
    // resident (persisting access window on this stream; opt-in, see phi_gpu_index_create)
    const bool l2_pin = ctx->l2_persist_bytes > 0;
    if (l2_pin) {
        cudaStreamAttrValue av; memset(&av, 0, sizeof av);
        av.accessPolicyWindow.base_ptr = ctx->coord.p;
        av.accessPolicyWindow.num_bytes = std::min<size_t>((size_t)V * 16, ctx->l2_window_max);
        av.accessPolicyWindow.hitRatio = std::min(1.0f, (float)ctx->l2_persist_bytes / (float)std::max<size_t>(av.accessPolicyWindow.num_bytes, 1));
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CUP(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    if (ctx->n_pieces && ctx->n_wpieces > 1) {
        // phi_gpu_index_run: the walk steps are still arriving; every piece is scanned as soon as it is there
        const uint64_t tile = walk_steps_fused_tile_steps();
        uint64_t t0 = 0;
        for (int p = 0; p < ctx->n_wpieces; ++p) {
            CUP(cudaStreamWaitEvent(st, ctx->ev_wpiece[p], 0));
            const uint64_t t1 = p + 1 == ctx->n_wpieces ? walk_steps_fused_tiles(S) : ctx->wpiece_end[p] / tile;
            if (t1 > t0 || p == 0)
                CUP(walk_steps_fused(d_walk_vtx, d_walk_off, H, S, ctx->coord.as<uint4>(), V, ctx->fs_state.as<unsigned long long>() + 1, ctx->fs_state.as<uint32_t>(),
                                    ctx->step_base.as<uint32_t>(), ctx->chunk_step.as<uint32_t>(), ctx->c_walk.as<uint32_t>(), ctx->walk_len.as<uint64_t>(), d_ctr,
                                    st, launches, t0, t1));
            t0 = std::max(t0, t1);
        }
    } else {
        if (ctx->n_pieces) CUP(cudaStreamWaitEvent(st, ctx->ev_wpiece[ctx->n_wpieces - 1], 0));
        CUP(walk_steps_fused(d_walk_vtx, d_walk_off, H, S, ctx->coord.as<uint4>(), V, ctx->fs_state.as<unsigned long long>() + 1, ctx->fs_state.as<uint32_t>(),
                            ctx->step_base.as<uint32_t>(), ctx->chunk_step.as<uint32_t>(), ctx->c_walk.as<uint32_t>(), ctx->walk_len.as<uint64_t>(), d_ctr,
                            st, launches));
    }
    if (l2_pin) {
        cudaStreamAttrValue av; memset(&av, 0, sizeof av);                 // window off again (num_bytes 0)
        CUP(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    CUP(read_counters_on(ctx, st, ctr, h_ctr));                                             // wait 1
    if (l2_pin) cudaCtxResetPersistingL2Cache();                          // the step kernel is done: its persisting lines become ordinary ones again
    if (h_ctr[CTR_BAD_VTX]) return ctx->fail2(PHI_ERR_ARG, "graph view: a walk step names a vertex id >= n_vtx");
    if (h_ctr[CTR_SEG_TOO_LONG]) return ctx->fail2(PHI_ERR_UNSUPPORTED, "segment of 2^31 bases or more");
    const bool fused = h_ctr[CTR_ZERO_STEPS] == 0;
    uint64_t last = 0; uint32_t NC = 0;
    // zero-length segments contribute no bases (ILP_index.cpp:364-381): the general path drops their steps and looks at the walks again
    if (!fused) { CUP(ctx->step_len.reserve(S * 4 + 4)); CUP(ctx->gbase.reserve((S + 1) * 8)); CUP(cudaMemsetAsync(d_ctr + CTR_NONMONO, 0, 8, st)); }
    for (int attempt = 0; !fused; ++attempt) {
        CUP(cudaMemsetAsync(d_ctr + CTR_ZERO_STEPS, 0, 8, st)); CUP(cudaMemsetAsync(d_ctr + CTR_CHUNK_FLAGS, 0, 8, st));
        CUP(walk_step_pass(d_walk_vtx, d_walk_off, H, S, ctx->coord.as<uint4>(), V, ctx->step_len.as<PackedStep>(), d_ctr, st, launches));
        CUP(scan_packed_steps(ctx->step_len.as<PackedStep>(), ctx->gbase.as<uint64_t>(), S, scan_scr.p, st, launches));
        CUP(cudaMemcpyAsync(&last, ctx->gbase.as<uint64_t>() + (S - 1), 8, cudaMemcpyDeviceToHost, st));
        CUP(read_counters_on(ctx, st, ctr, h_ctr));
        if (!h_ctr[CTR_ZERO_STEPS] || attempt) break;
        const uint64_t kept = S - h_ctr[CTR_ZERO_STEPS];
        CUP(flags.reserve(S * 4 + 4)); CUP(flags64.reserve((S + 1) * 8));
        CUP(ctx->walk_vtx_c.reserve(kept * 4 + 4)); CUP(ctx->walk_off_c.reserve(((size_t)H + 1) * 8));
        CUP(walk_compact_steps(d_walk_vtx, d_walk_off, H, S, ctx->step_len.as<PackedStep>(), flags.as<uint32_t>(), flags64.as<uint64_t>(),
                              scan_scr.p, kept, ctx->walk_vtx_c.as<uint32_t>(), ctx->walk_off_c.as<uint64_t>(), st, launches));
        d_walk_vtx = ctx->walk_vtx_c.as<uint32_t>(); d_walk_off = ctx->walk_off_c.as<uint64_t>(); n_steps_eff = S = kept;
        CUP(cudaMemsetAsync(d_ctr + CTR_NONMONO, 0, 8, st));
        if (!S) return PHI_OK;
    }
    (void)S0;
    walks_monotone = h_ctr[CTR_NONMONO] ? 0 : 1;
    if (h_ctr[CTR_CHUNK_FLAGS] >= (1ull << (64 - STEP_BASE_BITS)))
        return ctx->fail2(PHI_ERR_UNSUPPORTED, "too many walk chunks on one GPU: raise chunk_shift (phi_gpu_index_set_walk_sharing) or shard the walks");
    NC = (uint32_t)h_ctr[CTR_CHUNK_FLAGS];
    (void)last;
    ctx->n_chunks = NC;
    DevBuf *u32s[] = {&ctx->chunk_step, &ctx->c_walk, &ctx->c_L, &ctx->c_R, &ctx->c_lo, &ctx->c_hi, &ctx->c_slot, &ctx->c_rep, &ctx->c_ninst,
                      &ctx->c_ntile, &ctx->c_tile_base, &ctx->c_emitted, &ctx->c_hits, &ctx->c_surv, &ctx->c_surv_vtx};
    for (DevBuf *b : u32s) CUP(b->reserve(((size_t)NC + 2) * 4));
    CUP(ctx->c_h1.reserve(((size_t)NC + 1) * 8)); CUP(ctx->c_h2.reserve(((size_t)NC + 1) * 8));
    CUP(ctx->member_cnt.reserve(((size_t)NC + 2) * 4)); CUP(ctx->member_off.reserve(((size_t)NC + 2) * 8));
    ChunkTable C = chunk_table(ctx);
    if (!fused)
        CUP(walk_step_finalize(C, ctx->step_len.as<PackedStep>(), ctx->gbase.as<uint64_t>(), d_walk_off, H, S, ctx->step_base.as<uint32_t>(),
                              ctx->walk_len.as<uint64_t>(), st, launches));
    CUP(cudaMemcpyAsync(h_walk_len.data(), ctx->walk_len.p, (size_t)H * 8, cudaMemcpyDeviceToHost, st));
    CUP(chunk_keys(C, d_walk_vtx, d_walk_off, ctx->step_base.as<uint32_t>(), ctx->walk_len.as<uint64_t>(), ctx->coord.as<uint4>(), k, w, d_ctr, st, launches));
    uint32_t tcap = 1024; while (tcap < 2 * (uint64_t)NC) tcap <<= 1;
    CUP(ctx->ctable.reserve((size_t)tcap * 4));
    CUP(scan_scr.reserve(std::max(scan_u32_scratch((uint64_t)NC + 2), scan_u32_to_u64_scratch((uint64_t)NC + 2))));
    for (int dedupe = ctx->dedupe ? 1 : 0;; dedupe = 0) {
        CUP(cudaMemsetAsync(d_ctr + CTR_UNIQUE_WINDOWS, 0, 2 * 8, st));               // UNIQUE_WINDOWS, DEDUPE_MISMATCH
        CUP(chunk_group(C, ctx->ctable.as<uint32_t>(), tcap, d_walk_vtx, dedupe, w, d_ctr, st, launches));
        CUP(cudaMemsetAsync(ctx->c_ntile.as<uint32_t>() + NC, 0, 4, st));
        CUP(scan_u32(ctx->c_ntile.as<uint32_t>(), ctx->c_tile_base.as<uint32_t>(), (uint64_t)NC + 1, scan_scr.p, st, launches));
        uint32_t n_tiles = 0;
        CUP(cudaMemcpyAsync(&n_tiles, ctx->c_tile_base.as<uint32_t>() + NC, 4, cudaMemcpyDeviceToHost, st));
        CUP(read_counters_on(ctx, st, ctr, h_ctr));                                         // wait 2 (also: walk lengths)
        ctx->n_tiles = n_tiles;
        if (!dedupe || !h_ctr[CTR_DEDUPE_MISMATCH]) break;          // a fingerprint collision: sketch every chunk on its own
    }
    uint64_t total_bases = 0;
    for (uint32_t h = 0; h < H; ++h) {
        if (h_walk_len[h] >= (1ull << 31)) return ctx->fail2(PHI_ERR_UNSUPPORTED, "walk longer than 2^31-1 bases (the reference's int32 position loop overflows there too)");
        total_bases += h_walk_len[h];
    }
    if (total_bases >= (1ull << STEP_BASE_BITS)) return ctx->fail2(PHI_ERR_UNSUPPORTED, "2^38 or more walk bases on one GPU; shard the walks over more GPUs");
    ctx->unique_windows = h_ctr[CTR_UNIQUE_WINDOWS]; ctx->active_chunks = h_ctr[CTR_ACTIVE_CHUNKS]; ctx->path_pos = h_ctr[CTR_PATH_POS];
    CUP(ctx->tiles.reserve((size_t)ctx->n_tiles * sizeof(TileRec) + 32));
    CUP(chunk_tiles(C, d_walk_off, ctx->walk_len.as<uint64_t>(), ctx->step_base.as<uint32_t>(), w, ctx->tiles.as<TileRec>(), st, launches));
    // member walks of every representative (the grouped result copies them instead of instantiating one record per member)
    DevBuf *cms[] = {&ctx->fs_state, &ctx->cm_off, &ctx->cm_cursor, &ctx->cm_tmp, &ctx->cm_walk};
    for (DevBuf *b : cms) CUP(b->reserve(((size_t)NC + 2) * 4));
    CUP(chunk_members(C, H, ctx->cm_off.as<uint32_t>(), ctx->cm_cursor.as<uint32_t>(), ctx->cm_tmp.as<uint32_t>(), ctx->cm_walk.as<uint32_t>(),
                     scan_scr.p, st, launches));
    return PHI_OK;
}


// =====================================================================================================
// Multi-GPU: one ctx per GPU, NCCL for the two exchange steps of the path.
//   (1) distinct read-minimizer hashes -> owner GPU by hash range (all-to-all), owners dedup + sort, the sorted
//       slices are broadcast so every GPU holds the whole ranked spectrum (concatenation of range slices is sorted);
//   (2) walk hits -> owner of their rank (all-to-all), where the threshold filter and the final ordering run.
// The reference has no counterpart (single process, OpenMP); see DESIGN.md §6.
// =====================================================================================================
namespace {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclCommAbort) CommAbort = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
};

NcclApi *nccl_api(std::string &err)
{
    static NcclApi api; static std::mutex mu; static bool tried = false;
    std::lock_guard<std::mutex> lk(mu);
    if (api.handle) return &api;
    if (tried) { err = "NCCL could not be loaded"; return nullptr; }
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
    if (!api.handle) { err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return nullptr; }
#define PHI_NCCL_SYM(field, name) api.field = (decltype(api.field))dlsym(api.handle, name); if (!api.field) { err = std::string("NCCL symbol missing: ") + name; api.handle = nullptr; return nullptr; }
    PHI_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") PHI_NCCL_SYM(CommInitRank, "ncclCommInitRank") PHI_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    PHI_NCCL_SYM(GetErrorString, "ncclGetErrorString") PHI_NCCL_SYM(GroupStart, "ncclGroupStart") PHI_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    PHI_NCCL_SYM(Send, "ncclSend") PHI_NCCL_SYM(Recv, "ncclRecv") PHI_NCCL_SYM(AllGather, "ncclAllGather") PHI_NCCL_SYM(Broadcast, "ncclBroadcast")
    PHI_NCCL_SYM(CommAbort, "ncclCommAbort") PHI_NCCL_SYM(AllReduce, "ncclAllReduce")
#undef PHI_NCCL_SYM
    return &api;
}

}  // namespace

// the ctx's communicator goes away: orderly (destroy) or, after a failure in the middle of an exchange, by abort so that this
// rank does not sit in a half-issued collective
static void comm_release(phi_gpu_index_ctx *ctx, bool abort)
{
    if (!ctx->comm) return;
    std::string err;
    NcclApi *nc = nccl_api(err);
    if (nc) { if (abort) nc->CommAbort((ncclComm_t)ctx->comm); else nc->CommDestroy((ncclComm_t)ctx->comm); }
    ctx->comm = nullptr;
}

#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return ctx->fail(PHI_ERR_COMM, std::string(#call) + ": " + nc->GetErrorString(r_)); } while (0)

extern "C" int phi_gpu_index_comm_unique_id(uint8_t id[PHI_COMM_ID_BYTES])
{
    std::string err;
    NcclApi *nc = nccl_api(err);
    if (!nc || !id) { g_create_error = err; return PHI_ERR_COMM; }
    static_assert(PHI_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
    ncclUniqueId u;
    if (nc->GetUniqueId(&u) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return PHI_ERR_COMM; }
    memcpy(id, u.internal, PHI_COMM_ID_BYTES);
    return PHI_OK;
}

extern "C" int phi_gpu_index_comm_init(phi_gpu_index_ctx *ctx, int rank, int world, const uint8_t id[PHI_COMM_ID_BYTES],
                                       uint32_t walk_id_base, uint32_t n_walks_global)
{
    if (!ctx) return PHI_ERR_ARG;
    if (world < 1 || rank < 0 || rank >= world || !id) return ctx->fail(PHI_ERR_ARG, "bad rank/world/id");
    if (world > 64) return ctx->fail(PHI_ERR_UNSUPPORTED, "more than 64 ranks");
    comm_release(ctx, false);                                               // a repeated comm_init replaces the communicator
    ctx->rank = rank; ctx->world = world; ctx->walk_id_base = walk_id_base; ctx->n_walks_global = n_walks_global;
    if (world == 1) return PHI_OK;
    std::string err;
    NcclApi *nc = nccl_api(err);
    if (!nc) return ctx->fail(PHI_ERR_COMM, err);
    CU(cudaSetDevice(ctx->device));
    // the exchanges are point-to-point (grouped send / recv by owner): give NCCL more than its default couple of channels per peer,
    // unless the caller has set the knobs (tens of MB per peer move at a fraction of NVLink speed otherwise)
    setenv("NCCL_MIN_P2P_NCHANNELS", "8", 0);
    setenv("NCCL_MAX_P2P_NCHANNELS", "32", 0);
    ncclUniqueId u; memcpy(u.internal, id, PHI_COMM_ID_BYTES);
    ncclComm_t comm;
    NC(nc->CommInitRank(&comm, world, u, rank));
    ctx->comm = comm;
    return PHI_OK;
}

// ---- small collectives of the exchange steps.  Every rank contributes `n` u64 words that already sit on the device (d_send);
// the gathered words come back on the host (pinned).  ONE collective + ONE copy + ONE wait: counts never travel host -> device -> host.
// The last word of every rank's contribution is its status: 0, or the error code of a stage that failed on that rank, so that all
// ranks leave the exchange together instead of one rank returning while its peers sit in the next collective.
static int allgather_words(phi_gpu_index_ctx *ctx, NcclApi *nc, const uint64_t *d_send, size_t n, const uint64_t *&h_all)
{
    const size_t W = ctx->world;
    CU(ctx->ag_recv.reserve(n * W * 8 + 8));
    if (ctx->h_words_cap < n * W) {
        if (ctx->h_words) cudaFreeHost(ctx->h_words);
        ctx->h_words = nullptr; ctx->h_words_cap = 0;
        CU(cudaHostAlloc((void **)&ctx->h_words, n * W * 8 + 64, cudaHostAllocDefault));
        ctx->h_words_cap = n * W;
    }
    NC(nc->AllGather(d_send, ctx->ag_recv.p, n, ncclUint64, (ncclComm_t)ctx->comm, ctx->st));
    CU(cudaMemcpyAsync(ctx->h_words, ctx->ag_recv.p, n * W * 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    h_all = ctx->h_words;
    return PHI_OK;
}

// did any rank report a failure?  (word `n - 1` of every rank's contribution)
static int peers_status(phi_gpu_index_ctx *ctx, const uint64_t *h_all, size_t n, int own_status)
{
    for (int p = 0; p < ctx->world; ++p) {
        const uint64_t st = h_all[(size_t)p * n + n - 1];
        if (!st) continue;
        if (p == ctx->rank) return own_status ? own_status : PHI_ERR_COMM;    // this rank's own error text is already in place
        return ctx->fail(PHI_ERR_COMM, "rank " + std::to_string(p) + " of the multi-GPU run failed (status " + std::to_string(st) + "); this rank stops with it");
    }
    return PHI_OK;
}

__global__ void owner_split_kernel(const uint64_t *sorted, uint64_t n, int world, uint64_t status, uint64_t *split /* [world + 2] */)
{
    int o = threadIdx.x;
    if (o > world + 1) return;
    if (o == world + 1) { split[o] = status; return; }
    if (o == world) { split[o] = n; return; }
    // smallest hash owned by o is ceil(o * 2^64 / world)   (phi_shard_owner_of_hash)
    const uint64_t key = (uint64_t)((((unsigned __int128)o << 64) + world - 1) / world);
    uint64_t lo = 0, hi = n;                                            // first index with sorted[i] >= key
    while (lo < hi) { uint64_t m = (lo + hi) >> 1; if (sorted[m] < key) lo = m + 1; else hi = m; }
    split[o] = lo;
}

// the owner's slice is complete: (distinct keys incl. the key ~0, overflow flag, status) for the second small collective
__global__ void owner_tail_kernel(const uint32_t *tblk_total, const unsigned long long *ctr, uint64_t status, uint64_t *slice, uint64_t *words /* [3] */)
{
    uint64_t n = tblk_total ? *tblk_total : 0;
    if (tblk_total && ctr[CTR_HAS_MAXKEY]) { slice[n] = TABLE_EMPTY; ++n; }   // the key ~0 is the largest of the last owner's range
    words[0] = n; words[1] = tblk_total ? ctr[CTR_OVERFLOW] : 0; words[2] = status;
}

// Exchange of the locally distinct, locally sorted hashes (n_local of them in ctx->spec_a) -> ctx->spec_a = global sorted spectrum,
// ctx->own_off = first global rank of every owner.  status: error code of an earlier stage of this rank (all ranks stop together).
static int exchange_spectrum(phi_gpu_index_ctx *ctx, uint64_t n_local, uint64_t &n_spec, int status)
{
    std::string err; NcclApi *nc = nccl_api(err);
    if (!nc || !ctx->comm) return ctx->fail(PHI_ERR_COMM, "communicator not initialised");
    const int W = ctx->world, me = ctx->rank;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    CU(ctx->xcnt.reserve(8192));
    uint64_t *d_split = ctx->xcnt.as<uint64_t>();                       // [W + 2] | tail words [3]
    uint64_t *d_tail = d_split + 72;
    owner_split_kernel<<<1, 128, 0, ctx->st>>>(ctx->spec_a.as<uint64_t>(), n_local, W, (uint64_t)status, d_split);
    CU(cudaGetLastError()); ctx->launches++;
    const uint64_t *all = nullptr;
    int rc = allgather_words(ctx, nc, d_split, W + 2, all);            // all[src * (W + 2) + o] = first index of owner o's part in src's array
    if (rc) return rc;
    if ((rc = peers_status(ctx, all, W + 2, status))) return rc;
    std::vector<uint64_t> scnt(W), soff(W), rcnt(W), roff(W);
    uint64_t rtot = 0;
    for (int p = 0; p < W; ++p) {
        const uint64_t *sp = all + (size_t)p * (W + 2), *mine = all + (size_t)me * (W + 2);
        scnt[p] = mine[p + 1] - mine[p]; soff[p] = mine[p];
        rcnt[p] = sp[me + 1] - sp[me]; roff[p] = rtot; rtot += rcnt[p];
    }
    // from here on a failure of this rank alone would leave the peers inside a collective: the communicator is aborted then
    struct AbortGuard { phi_gpu_index_ctx *c; bool armed; ~AbortGuard() { if (armed) comm_release(c, true); } } guard = {ctx, true};
    // xk_b will hold this owner's slice and is read up to the LARGEST owner's slice size by the all-gather below: no owner can
    // end up with more distinct keys than it was sent, so the largest receive count of any owner bounds them all
    uint64_t rtot_max = 0;
    for (int o = 0; o < W; ++o) { uint64_t t = 0; for (int p = 0; p < W; ++p) t += all[(size_t)p * (W + 2) + o + 1] - all[(size_t)p * (W + 2) + o]; rtot_max = std::max(rtot_max, t); }
    CU(ctx->xk_a.reserve((rtot + 1) * 8)); CU(ctx->xk_b.reserve((rtot_max + 2) * 8));
    for (uint64_t cap_mul = 2;; cap_mul <<= 1) {
        NC(nc->GroupStart());
        for (int p = 0; p < W; ++p) {
            if (p == me) continue;
            if (scnt[p]) NC(nc->Send(ctx->spec_a.as<uint64_t>() + soff[p], scnt[p], ncclUint64, p, comm, ctx->st));
            if (rcnt[p]) NC(nc->Recv(ctx->xk_a.as<uint64_t>() + roff[p], rcnt[p], ncclUint64, p, comm, ctx->st));
        }
        NC(nc->GroupEnd());
        if (scnt[me]) CU(cudaMemcpyAsync(ctx->xk_a.as<uint64_t>() + roff[me], ctx->spec_a.as<uint64_t>() + soff[me], scnt[me] * 8, cudaMemcpyDeviceToDevice, ctx->st));
        // owner: what arrived goes into an order-preserving table over this owner's hash range (duplicates collapse), the probe
        // clusters are sorted in place and the occupied slots are written out in order: the owner's sorted distinct slice
        unsigned long long *d_ctr = ctx->ctr.as<unsigned long long>();
        const uint32_t *d_total = nullptr;
        if (rtot) {
            const uint64_t base = (uint64_t)((((unsigned __int128)me << 64) + W - 1) / W);
            const unsigned __int128 range = (me + 1 < W ? (unsigned __int128)(uint64_t)((((unsigned __int128)(me + 1) << 64) + W - 1) / W) : ((unsigned __int128)1 << 64)) - base;
            uint64_t cap = 1024; while (cap < cap_mul * rtot) cap <<= 1;
            const uint64_t limit = cap + TABLE_PAD;
            const uint64_t mult = (uint64_t)((((unsigned __int128)cap) << 64) / range);    // home = umulhi(key - base, mult) < cap
            const size_t nb = table_blocks(limit);
            CU(ctx->table.reserve((limit + 1) * 8)); CU(ctx->tblk.reserve((nb + 2) * 4));
            CU(ctx->scan_scr.reserve(std::max(scan_u32_scratch(nb + 1), (size_t)1024)));
            CU(fill_u64(ctx->table.as<uint64_t>(), limit + 1, TABLE_EMPTY, ctx->st, &ctx->launches));
            CU(cudaMemsetAsync(d_ctr + CTR_OVERFLOW, 0, 2 * 8, ctx->st));                   // OVERFLOW, HAS_MAXKEY
            CU(table_insert_keys(ctx->xk_a.as<uint64_t>(), rtot, ctx->table.as<uint64_t>(), base, mult, limit, d_ctr, ctx->st, &ctx->launches));
            CU(table_sort_and_count(ctx->table.as<uint64_t>(), limit, ctx->tblk.as<uint32_t>(), ctx->st, &ctx->launches));
            CU(cudaMemsetAsync(ctx->tblk.as<uint32_t>() + nb, 0, 4, ctx->st));
            CU(scan_u32_inplace(ctx->tblk.as<uint32_t>(), nb + 1, ctx->scan_scr.p, ctx->st, &ctx->launches));
            CU(table_write_ordered(ctx->table.as<uint64_t>(), limit, ctx->tblk.as<uint32_t>(), ctx->xk_b.as<uint64_t>(), ctx->st, &ctx->launches));
            d_total = ctx->tblk.as<uint32_t>() + nb;
        }
        owner_tail_kernel<<<1, 1, 0, ctx->st>>>(d_total, d_ctr, 0, ctx->xk_b.as<uint64_t>(), d_tail);
        CU(cudaGetLastError()); ctx->launches++;
        rc = allgather_words(ctx, nc, d_tail, 3, all);                  // (distinct keys of the owner, overflow, status) of every owner
        if (rc) return rc;
        bool overflow = false;
        for (int p = 0; p < W; ++p) overflow |= all[(size_t)p * 3 + 1] != 0;
        if (!overflow) break;                                           // (every rank sees the same flags and repeats, or not, together)
        if (cap_mul > 64) return ctx->fail(PHI_ERR_CUDA, "owner-side spectrum table overflowed repeatedly (internal error)");
    }
    ctx->own_off.assign(W + 1, 0);
    for (int o = 0; o < W; ++o) ctx->own_off[o + 1] = ctx->own_off[o] + all[(size_t)o * 3];
    n_spec = ctx->own_off[W];
    CU(ctx->spec_a.reserve((n_spec + 1) * 8));
    // every owner's sorted slice to everybody: concatenation of range slices is the sorted spectrum.  ONE all-gather of slices padded
    // to the largest one (Murmur hashes are uniform: the owners' slices differ by a fraction of a per cent) — a collective NCCL runs
    // at full NVLink / NVSwitch bandwidth, unlike W - 1 point-to-point pairs — then the slices are moved next to each other.
    uint64_t maxn = 0;
    for (int o = 0; o < W; ++o) maxn = std::max(maxn, ctx->own_off[o + 1] - ctx->own_off[o]);
    if (maxn) {
        CU(ctx->xk_a.reserve((size_t)W * maxn * 8 + 8));                  // (the keys received above are not needed any more)
        NC(nc->AllGather(ctx->xk_b.p, ctx->xk_a.p, maxn, ncclUint64, comm, ctx->st));
        for (int o = 0; o < W; ++o) {
            const uint64_t cnt = ctx->own_off[o + 1] - ctx->own_off[o];
            if (cnt) CU(cudaMemcpyAsync(ctx->spec_a.as<uint64_t>() + ctx->own_off[o], ctx->xk_a.as<uint64_t>() + (size_t)o * maxn, cnt * 8, cudaMemcpyDeviceToDevice, ctx->st));
        }
    }
    guard.armed = false;
    return PHI_OK;
}

// ---- record routing (group summaries) to the owner of their rank.  A record is (rank, count, vertex list).  Every (src -> dst)
// pair moves ONE packed byte stream:  voff u64[n] | rank u32[n] | cnt u32[n] | nv u8[n] (padded to 8) | vtx i32[nvtx]
constexpr int ROUTE_ITEMS = 8;                      // records per thread: 2048 per block -> few global atomics
struct RouteIn {
    const uint32_t *rank, *cnt; const uint64_t *voff; const uint8_t *nv; const int32_t *vtx;
    uint64_t n;
};
__host__ __device__ inline uint64_t route_seg_bytes(uint64_t n, uint64_t nvtx) { return 16 * n + ((n + 7) & ~7ull) + ((4 * nvtx + 7) & ~7ull); }
__device__ __forceinline__ int owner_of_rank(const uint64_t *own_off, int world, uint64_t r)
{
    int o = 0;
    while (o + 1 < world && own_off[o + 1] <= r) ++o;
    return o;
}
__global__ void __launch_bounds__(256) route_count_kernel(RouteIn I, const uint64_t *own_off, int world, unsigned long long *cnt /* [2*world] */)
{
    __shared__ unsigned long long sh[128];
    for (int i = threadIdx.x; i < 2 * world; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * (256 * ROUTE_ITEMS);
    #pragma unroll
    for (int it = 0; it < ROUTE_ITEMS; ++it) {
        uint64_t i = base + it * 256 + threadIdx.x;
        int o = -1; uint32_t nv = 0;
        if (i < I.n) { o = owner_of_rank(own_off, world, I.rank[i]); nv = I.nv[i]; }
        // warp-aggregated: one shared-memory atomic per distinct owner per warp
        uint32_t peers = __match_any_sync(0xFFFFFFFFu, o);
        if (o >= 0) { atomicAdd(&sh[world + o], (unsigned long long)nv); if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sh[o], (unsigned long long)__popc(peers)); }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * world; j += blockDim.x) if (sh[j]) atomicAdd(&cnt[j], sh[j]);
}

struct RouteSeg { uint64_t byte_off, n, nvtx; };    // one (src -> dst) segment inside a packed buffer
struct RouteLayout { RouteSeg seg[64]; };
__device__ __forceinline__ void route_seg_ptrs(unsigned char *buf, const RouteSeg &s, uint64_t *&voff, uint32_t *&rank, uint32_t *&cnt, uint8_t *&nv, int32_t *&vtx)
{
    unsigned char *p = buf + s.byte_off;
    voff = (uint64_t *)p; rank = (uint32_t *)(p + 8 * s.n); cnt = rank + s.n; nv = (uint8_t *)(cnt + s.n); vtx = (int32_t *)(nv + ((s.n + 7) & ~7ull));
}
__global__ void __launch_bounds__(256) route_pack_kernel(RouteIn I, RouteLayout L, unsigned char *buf, unsigned long long *cursor /* [2*world], zeroed */,
                                                         const uint64_t *own_off, int world)
{
    __shared__ unsigned long long sh_cnt[128], sh_base[128];
    for (int i = threadIdx.x; i < 2 * world; i += blockDim.x) sh_cnt[i] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * (256 * ROUTE_ITEMS);
    int own[ROUTE_ITEMS]; uint32_t lh[ROUTE_ITEMS], lv[ROUTE_ITEMS];
    #pragma unroll
    for (int it = 0; it < ROUTE_ITEMS; ++it) {
        uint64_t i = base + it * 256 + threadIdx.x;
        own[it] = -1; lh[it] = lv[it] = 0;
        if (i < I.n) {
            int o = own[it] = owner_of_rank(own_off, world, I.rank[i]);
            lh[it] = (uint32_t)atomicAdd(&sh_cnt[o], 1ull); lv[it] = (uint32_t)atomicAdd(&sh_cnt[world + o], (unsigned long long)I.nv[i]);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * world; j += blockDim.x) sh_base[j] = sh_cnt[j] ? atomicAdd(&cursor[j], sh_cnt[j]) : 0ull;
    __syncthreads();
    #pragma unroll
    for (int it = 0; it < ROUTE_ITEMS; ++it) {
        if (own[it] < 0) continue;
        const uint64_t i = base + it * 256 + threadIdx.x;
        const int o = own[it];
        uint64_t *s_voff; uint32_t *s_rank, *s_cnt; uint8_t *s_nv; int32_t *s_vtx;
        route_seg_ptrs(buf, L.seg[o], s_voff, s_rank, s_cnt, s_nv, s_vtx);
        const unsigned long long dh = sh_base[o] + lh[it], dv = sh_base[world + o] + lv[it];
        const uint32_t nv = I.nv[i];
        s_rank[dh] = I.rank[i]; s_cnt[dh] = I.cnt[i]; s_nv[dh] = (uint8_t)nv; s_voff[dh] = dv;   // vertex offset inside the segment
        const int32_t *src = I.vtx + I.voff[i];
        for (uint32_t q = 0; q < nv; ++q) s_vtx[dv + q] = src[q];
    }
}
// received segments -> flat record arrays (the owner-side group table indexes records, not segments)
__global__ void __launch_bounds__(256) route_unpack_kernel(RouteLayout L, int world, unsigned char *buf, const uint64_t rec0[65], const uint64_t vtx0[65],
                                                           uint32_t *r_rank, uint32_t *r_cnt, uint64_t *r_voff, uint8_t *r_nv, int32_t *r_vtx)
{
    const int p = blockIdx.y;
    const RouteSeg s = L.seg[p];
    uint64_t *s_voff; uint32_t *s_rank, *s_cnt; uint8_t *s_nv; int32_t *s_vtx;
    route_seg_ptrs(buf, s, s_voff, s_rank, s_cnt, s_nv, s_vtx);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (uint64_t i = t0; i < s.n; i += stride) {
        const uint64_t d = rec0[p] + i;
        r_rank[d] = s_rank[i]; r_cnt[d] = s_cnt[i]; r_nv[d] = s_nv[i]; r_voff[d] = vtx0[p] + s_voff[i];
    }
    for (uint64_t i = t0; i < s.nvtx; i += stride) r_vtx[vtx0[p] + i] = s_vtx[i];
}

// Route the records of `I` to the owners of their ranks.  Received records land in ctx->r_rank / r_walk (= count) / r_voff / r_nv /
// r_vtx (rh records, rv vertices).  status: error code of an earlier stage of this rank (all ranks stop together).
static int exchange_records(phi_gpu_index_ctx *ctx, const RouteIn &I, uint64_t &rh, uint64_t &rv, int status)
{
    std::string err; NcclApi *nc = nccl_api(err);
    if (!nc || !ctx->comm) return ctx->fail(PHI_ERR_COMM, "communicator not initialised");
    const int W = ctx->world, me = ctx->rank;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    const uint64_t n = I.n;
    const unsigned nblk = (unsigned)((n + 256 * ROUTE_ITEMS - 1) / (256 * ROUTE_ITEMS));
    CU(ctx->xcnt.reserve(8192));
    uint64_t *d_own = ctx->xcnt.as<uint64_t>() + 128;                   // [65] own_off | [2W + 1] counts + status | [2W] cursors | [65] rec0 | [65] vtx0
    unsigned long long *d_cnt = (unsigned long long *)(d_own + 65), *d_cur = d_cnt + 130;
    uint64_t *d_rec0 = (uint64_t *)(d_cur + 128), *d_vtx0 = d_rec0 + 65;
    // own_off, zeroed counts and the status word in one small copy from pinned memory
    uint64_t *hp = ctx->h_route;
    for (int o = 0; o <= W; ++o) hp[o] = ctx->own_off[o];
    for (int j = 0; j < 2 * W; ++j) hp[65 + j] = 0;
    hp[65 + 2 * W] = (uint64_t)status;
    CU(cudaMemcpyAsync(d_own, hp, (65 + 2 * W + 1) * 8, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemsetAsync(d_cur, 0, 2 * W * 8, ctx->st));
    if (n) {
        route_count_kernel<<<nblk, 256, 0, ctx->st>>>(I, d_own, W, d_cnt);
        CU(cudaGetLastError()); ctx->launches++;
    }
    const uint64_t *all = nullptr;
    const size_t NWORDS = 2 * W + 1;
    int rc = allgather_words(ctx, nc, (const uint64_t *)d_cnt, NWORDS, all);   // all[src * NWORDS + {dst, W + dst}] = records, vertices
    if (rc) return rc;
    if ((rc = peers_status(ctx, all, NWORDS, status))) return rc;
    RouteLayout SL, RL; memset(&SL, 0, sizeof SL); memset(&RL, 0, sizeof RL);
    uint64_t sbytes = 0, rbytes = 0; rh = rv = 0;
    uint64_t *h_rec0 = hp + 256, *h_vtx0 = hp + 256 + 65;
    for (int p = 0; p < W; ++p) {
        const uint64_t *mine = all + (size_t)me * NWORDS, *src = all + (size_t)p * NWORDS;
        SL.seg[p] = {sbytes, mine[p], mine[W + p]}; sbytes += route_seg_bytes(mine[p], mine[W + p]);
        RL.seg[p] = {rbytes, src[me], src[W + me]}; rbytes += route_seg_bytes(src[me], src[W + me]);
        h_rec0[p] = rh; h_vtx0[p] = rv; rh += src[me]; rv += src[W + me];
    }
    if (rh >= (1ull << 32)) return ctx->fail(PHI_ERR_UNSUPPORTED, "more than 2^32-1 records routed to one GPU");
    struct AbortGuard { phi_gpu_index_ctx *c; bool armed; ~AbortGuard() { if (armed) comm_release(c, true); } } guard = {ctx, true};
    CU(ctx->s_vtx.reserve(sbytes + 64)); CU(ctx->s_voff.reserve(rbytes + 64));           // packed send / receive buffers
    CU(ctx->r_rank.reserve(rh * 4 + 4)); CU(ctx->r_walk.reserve(rh * 4 + 4)); CU(ctx->r_voff.reserve(rh * 8 + 8)); CU(ctx->r_nv.reserve(rh + 4)); CU(ctx->r_vtx.reserve(rv * 4 + 4));
    unsigned char *sbuf = ctx->s_vtx.as<unsigned char>(), *rbuf = ctx->s_voff.as<unsigned char>();
    if (n) {
        route_pack_kernel<<<nblk, 256, 0, ctx->st>>>(I, SL, sbuf, d_cur, d_own, W);
        CU(cudaGetLastError()); ctx->launches++;
    }
    NC(nc->GroupStart());
    for (int p = 0; p < W; ++p) {
        if (p == me) continue;
        const uint64_t sb = route_seg_bytes(SL.seg[p].n, SL.seg[p].nvtx), rb = route_seg_bytes(RL.seg[p].n, RL.seg[p].nvtx);
        if (SL.seg[p].n) NC(nc->Send(sbuf + SL.seg[p].byte_off, sb, ncclUint8, p, comm, ctx->st));
        if (RL.seg[p].n) NC(nc->Recv(rbuf + RL.seg[p].byte_off, rb, ncclUint8, p, comm, ctx->st));
    }
    NC(nc->GroupEnd());
    if (SL.seg[me].n) CU(cudaMemcpyAsync(rbuf + RL.seg[me].byte_off, sbuf + SL.seg[me].byte_off, route_seg_bytes(SL.seg[me].n, SL.seg[me].nvtx), cudaMemcpyDeviceToDevice, ctx->st));
    if (rh) {
        CU(cudaMemcpyAsync(d_rec0, h_rec0, 130 * 8, cudaMemcpyHostToDevice, ctx->st));   // rec0 | vtx0 (pinned: stays valid until the next exchange)
        uint64_t big = 0; for (int p = 0; p < W; ++p) big = std::max(big, std::max(RL.seg[p].n, RL.seg[p].nvtx));
        dim3 grid((unsigned)std::min<uint64_t>((big + 255) / 256, 4096), (unsigned)W);
        route_unpack_kernel<<<grid, 256, 0, ctx->st>>>(RL, W, rbuf, d_rec0, d_vtx0, ctx->r_rank.as<uint32_t>(), ctx->r_walk.as<uint32_t>(), ctx->r_voff.as<uint64_t>(),
                                                      ctx->r_nv.as<uint8_t>(), ctx->r_vtx.as<int32_t>());
        CU(cudaGetLastError()); ctx->launches++;
    }
    guard.armed = false;
    return PHI_OK;
}

// ---- stage: reads -> ranked spectrum (sorted distinct hashes + radix directory).  Two halves so that the graph
// preparation can be issued on the second stream while the read kernel runs.
namespace { struct ReadsState { uint64_t n_tiles = 0, cap = 0; bool relaunch = false; }; }

static int reads_sketch_launch(phi_gpu_index_ctx *ctx, int k, int w, const ReadsState &rs)
{
    unsigned long long *d_ctr = ctx->ctr.as<unsigned long long>();
    const uint64_t limit = rs.cap + TABLE_PAD;                              // slots [0, limit) + the EMPTY sentinel
    CU(ctx->table.reserve((limit + 1) * 8));
    CU(fill_u64(ctx->table.as<uint64_t>(), limit + 1, TABLE_EMPTY, ctx->st, &ctx->launches));
    CU(cudaMemsetAsync(d_ctr, 0, 4 * 8, ctx->st));                         // DISTINCT (unused), OVERFLOW, HAS_MAXKEY, READ_EMITTED
    ReadSketchArgs A;
    A.layout = read_tile_layout(k, w);
    A.read_bases = ctx->read_bases.as<uint8_t>() + 16; A.read_off = ctx->read_off.as<uint64_t>();
    A.n_reads = ctx->n_reads; A.total_bases = ctx->read_total; A.tile_first_read = ctx->tile_first_read.as<uint64_t>();
    A.k = k; A.w = w; A.table = ctx->table.as<uint64_t>(); A.table_mult = rs.cap; A.table_limit = limit; A.ctr = d_ctr;
    { static int bulk = -1; if (bulk < 0) { const char *e = getenv("PHI_GPU_READ_BULK"); bulk = e ? atoi(e) : 1; } A.bulk = bulk; }
    CU(cudaEventRecord(ctx->ev[EV_RK0], ctx->st));
    if (ctx->n_pieces > 1 && !rs.relaunch) {
        // the reads are still arriving: sketch the tiles whose bases are complete after every piece
        uint64_t t0 = 0;
        for (int p = 0; p < ctx->n_pieces; ++p) {
            CU(cudaStreamWaitEvent(ctx->st, ctx->ev_piece[p], 0));
            uint64_t t1 = rs.n_tiles;
            if (p + 1 < ctx->n_pieces) {                                   // last tile with  tile*cap - w - pad + NB <= piece_end
                const long long lim = (long long)ctx->piece_end[p] + A.layout.pad + w - A.layout.NB;
                t1 = lim < 0 ? 0 : std::min<uint64_t>(rs.n_tiles, (uint64_t)lim / (uint64_t)A.layout.cap + 1);
            }
            if (t1 > t0) { A.tile0 = t0; CU(launch_read_sketch(A, t1 - t0, ctx->st)); ctx->launches++; t0 = t1; }
        }
    } else {
        if (ctx->n_pieces) CU(cudaStreamWaitEvent(ctx->st, ctx->ev_piece[ctx->n_pieces - 1], 0));
        A.tile0 = 0;
        CU(launch_read_sketch(A, rs.n_tiles, ctx->st)); ctx->launches++;
    }
    CU(cudaEventRecord(ctx->ev[EV_RK1], ctx->st));
    // probe clusters sorted in place -> the table is in ascending order; per-block counts -> offsets -> total
    const size_t nb = table_blocks(limit);
    CU(ctx->tblk.reserve((nb + 2) * 4));
    CU(ctx->scan_scr.reserve(std::max(scan_u32_scratch(nb + 1), (size_t)1024)));
    CU(table_sort_and_count(ctx->table.as<uint64_t>(), limit, ctx->tblk.as<uint32_t>(), ctx->st, &ctx->launches));
    CU(cudaMemsetAsync(ctx->tblk.as<uint32_t>() + nb, 0, 4, ctx->st));
    CU(scan_u32_inplace(ctx->tblk.as<uint32_t>(), nb + 1, ctx->scan_scr.p, ctx->st, &ctx->launches));
    CU(cudaMemcpyAsync(ctx->h_tot, ctx->tblk.as<uint32_t>() + nb, 4, cudaMemcpyDeviceToHost, ctx->st));
    return PHI_OK;
}

static int stage_reads_begin(phi_gpu_index_ctx *ctx, int k, int w, ReadsState &rs)
{
    unsigned long long *d_ctr = ctx->ctr.as<unsigned long long>();
    const int T = read_tile_windows(w);
    const uint64_t G = ctx->read_total, R = ctx->n_reads;
    rs.n_tiles = (R && G >= (uint64_t)(w + k - 1)) ? (G - k) / T + 1 : 0;
    CU(cudaEventRecord(ctx->ev[EV_RD0], ctx->st));
    CU(cudaEventRecord(ctx->ev[EV_RK0], ctx->st));
    CU(cudaEventRecord(ctx->ev[EV_RK1], ctx->st));
    if (ctx->n_pieces) CU(cudaStreamWaitEvent(ctx->st, ctx->ev_piece[0], 0));   // read offsets (and the first piece) have arrived
    if (!rs.n_tiles) return PHI_OK;
    count_positions_kernel<<<(unsigned)((R + 255) / 256), 256, 0, ctx->st>>>(ctx->read_off.as<uint64_t>(), R, k, w, d_ctr + CTR_READ_POS);
    CU(cudaGetLastError()); ctx->launches++;
    CU(ctx->tile_first_read.reserve(rs.n_tiles * 8));
    CU(launch_read_tile_dir(ctx->read_off.as<uint64_t>(), R, w, rs.n_tiles, ctx->tile_first_read.as<uint64_t>(), ctx->st)); ctx->launches++;
    // expected minimizer density is 2/(w+1); size the table for ~35% load at that density, retry on overflow.  Reads repeat k-mers
    // (coverage), so far fewer keys are DISTINCT than emitted: when this ctx has sketched a read set of about this size before, the
    // table is sized from the distinct count seen then (every pass over the table streams all of its slots, occupied or not).
    double dens = std::min(1.0, 2.6 / (w + 1.0));
    uint64_t want = (uint64_t)((double)G * dens * 2.0) + 1024;
    if (ctx->spec_hint && G <= ctx->spec_hint_bases + ctx->spec_hint_bases / 8 && G + G / 8 >= ctx->spec_hint_bases)
        want = std::min(want, ctx->spec_hint * 5 / 2 + 1024);
    rs.cap = 1024;
    while (rs.cap < want) rs.cap <<= 1;
    return reads_sketch_launch(ctx, k, w, rs);
}

static int stage_reads_local(phi_gpu_index_ctx *ctx, int k, int w, ReadsState &rs, RunOut &o, uint64_t &n_spec)
{
    n_spec = 0;
    if (rs.n_tiles) {
        for (;;) {
            CU(read_counters(ctx));
            if (!ctx->h_ctr[CTR_OVERFLOW]) break;                          // probing ran off the padding (table far too small): double it
            rs.cap <<= 1; rs.relaunch = true;
            int rc = reads_sketch_launch(ctx, k, w, rs);
            if (rc) return rc;
        }
        const uint64_t cap = rs.cap;
        o.read_emitted = ctx->h_ctr[CTR_READ_EMITTED];
        o.read_pos = ctx->h_ctr[CTR_READ_POS];
        const uint64_t nd = *ctx->h_tot;                                   // occupied slots (copied by the same sync)
        ctx->spec_hint = nd; ctx->spec_hint_bases = ctx->read_total;
        const bool maxkey = ctx->h_ctr[CTR_HAS_MAXKEY] != 0;
        n_spec = nd + (maxkey ? 1 : 0);
        if (n_spec >= (1ull << 31)) return ctx->fail(PHI_ERR_UNSUPPORTED, "more than 2^31-1 distinct read minimizers (count_sp_r is int32 in the reference)");
        CU(cudaEventRecord(ctx->ev[EV_READS], ctx->st));
        CU(ctx->spec_a.reserve((n_spec + 1) * 8));
        CU(ctx->spec_b.reserve((n_spec + 1) * 8));
        CU(ctx->sort_scr.reserve(radix_sort_scratch(std::max<uint64_t>(n_spec, 1))));
        CU(table_write_ordered(ctx->table.as<uint64_t>(), cap + TABLE_PAD, ctx->tblk.as<uint32_t>(), ctx->spec_a.as<uint64_t>(), ctx->st, &ctx->launches));
        if (maxkey) CU(fill_u64(ctx->spec_a.as<uint64_t>() + nd, 1, TABLE_EMPTY, ctx->st, &ctx->launches));
    } else {
        CU(cudaEventRecord(ctx->ev[EV_READS], ctx->st));
        CU(ctx->spec_a.reserve(8));
    }
    return PHI_OK;
}

// status (several GPUs only): error code of an earlier stage of this rank; it travels with the first small collective of the
// exchange so that all ranks stop together.
static int stage_reads_finish(phi_gpu_index_ctx *ctx, int k, int w, ReadsState &rs, RunOut &o, int &dbits, int status)
{
    uint64_t n_spec = 0;
    int rc_local = status ? status : stage_reads_local(ctx, k, w, rs, o, n_spec);
    if (rc_local && ctx->world == 1) return rc_local;
    if (ctx->world > 1) {                                                 // every rank takes part, also with zero local reads or a failure to report
        CU(cudaEventRecord(ctx->ev[EV_XS0], ctx->st));
        int rc = exchange_spectrum(ctx, rc_local ? 0 : n_spec, n_spec, rc_local);
        if (rc) return rc;
        CU(cudaEventRecord(ctx->ev[EV_XS1], ctx->st));
        if (n_spec >= (1ull << 31)) return ctx->fail(PHI_ERR_UNSUPPORTED, "more than 2^31-1 distinct read minimizers (count_sp_r is int32 in the reference)");
    }
    o.n_spec = (uint32_t)n_spec;
    dbits = 0;
    while (dbits < 30 && (1ull << dbits) < n_spec) ++dbits;
    CU(ctx->dir.reserve(((1ull << dbits) + 2) * 4));
    CU(build_directory(ctx->spec_a.as<uint64_t>(), o.n_spec, dbits, ctx->dir.as<uint32_t>(), ctx->st, &ctx->launches));
    return PHI_OK;
}

// ---- stage: representative chunks -> hits
static int stage_walks(phi_gpu_index_ctx *ctx, int k, int w, int mode, int dbits, const std::vector<uint64_t> &h_walk_len,
                       const uint32_t *d_walk_vtx, const uint64_t *d_walk_off, uint64_t n_steps_eff, int walks_monotone, RunOut &o)
{
    const uint32_t H = ctx->n_walks;
    const uint32_t HM = ctx->world > 1 ? ctx->n_walks_global : H;         // per-walk counters are indexed by global walk id
    const uint32_t wbase = ctx->world > 1 ? ctx->walk_id_base : 0;
    unsigned long long *d_ctr = ctx->ctr.as<unsigned long long>();
    CU(ctx->mpw.reserve(((size_t)HM + 1) * 8));
    uint64_t bases = 0;
    for (uint32_t h = 0; h < H; ++h) bases += h_walk_len[h];
    o.path_pos = ctx->path_pos;                                          // counted with the chunk geometry: only owned chunks when a walk region is set
    CU(cudaEventRecord(ctx->ev[EV_WK0], ctx->st));
    CU(cudaEventRecord(ctx->ev[EV_WK1], ctx->st));
    CU(cudaMemsetAsync(ctx->mpw.p, 0, ((size_t)HM + 1) * 8, ctx->st));
    o.n_hits = o.n_hit_vtx = 0; o.path_hits = 0;
    const uint32_t NT_ = ctx->n_tiles, NC = ctx->n_chunks;
    if (!H || !NT_) return PHI_OK;
    // capacity estimate: emitted density 2/(w+1), vertices per anchor 1 + (k-1)/mean node length; exact re-run on overflow
    double dens = std::min(1.0, 2.0 / (w + 1.0)) * 1.3;
    double mean_node = n_steps_eff ? (double)bases / (double)n_steps_eff : 1.0;
    uint64_t hit_cap = (uint64_t)((double)ctx->unique_windows * dens) + 65536;
    uint64_t vtx_cap = (uint64_t)((double)hit_cap * std::min((double)k, 1.0 + (k - 1) / std::max(mean_node, 1.0)) * 1.2) + 65536;
    CU(ctx->hseg_off.reserve((size_t)NT_ * SEG_PER_TILE * 4)); CU(ctx->hseg_cnt.reserve((size_t)NT_ * SEG_PER_TILE * 4));
    for (int attempt = 0;; ++attempt) {
        if (hit_cap >= (1ull << 32)) return ctx->fail(PHI_ERR_UNSUPPORTED, "more than 2^32-1 walk hits on one GPU; shard the walks over more GPUs");
        CU(ctx->hit_rank.reserve(hit_cap * 4)); CU(ctx->hit_chunk.reserve(hit_cap * 4)); CU(ctx->hit_pos.reserve(hit_cap * 4));
        CU(ctx->hit_voff.reserve(hit_cap * 8)); CU(ctx->hit_nv.reserve(hit_cap));
        if (mode == WALK_MODE_ALL) CU(ctx->hit_hash.reserve(hit_cap * 8));
        CU(ctx->vtx_pool.reserve(vtx_cap * 4));
        if (mode == WALK_MODE_PROBE) CU(ctx->probe.reserve(hit_cap * 32 + 32));
        CU(cudaMemsetAsync(d_ctr + CTR_HITS, 0, 2 * 8, ctx->st));
        CU(cudaMemsetAsync(ctx->hseg_cnt.p, 0, (size_t)NT_ * SEG_PER_TILE * 4, ctx->st));
        CU(cudaMemsetAsync(ctx->c_emitted.p, 0, (size_t)NC * 4, ctx->st));
        CU(cudaMemsetAsync(ctx->c_hits.p, 0, (size_t)NC * 4, ctx->st));
        WalkSketchArgs A;
        A.layout = walk_tile_layout(k, w); A.walks_monotone = walks_monotone;
        A.seg_bases = ctx->seg_bases.as<uint8_t>() + 16; A.seg_off = ctx->seg_off.as<uint64_t>(); A.top_order_map = ctx->top_order.as<int32_t>();
        A.walk_vtx = d_walk_vtx; A.walk_off = d_walk_off; A.step_base = ctx->step_base.as<uint32_t>();
        A.walk_len = ctx->walk_len.as<uint64_t>(); A.tiles = ctx->tiles.as<TileRec>();
        A.k = k; A.w = w; A.mode = mode;
        A.spec = ctx->spec_a.as<uint64_t>(); A.dir = ctx->dir.as<uint32_t>(); A.dbits = dbits;
        A.chunk_emitted = ctx->c_emitted.as<uint32_t>(); A.chunk_hits = ctx->c_hits.as<uint32_t>();
        A.hseg_off = ctx->hseg_off.as<uint32_t>(); A.hseg_cnt = ctx->hseg_cnt.as<uint32_t>();
        A.hit_rank = ctx->hit_rank.as<uint32_t>(); A.hit_chunk = ctx->hit_chunk.as<uint32_t>(); A.hit_pos = ctx->hit_pos.as<uint32_t>();
        A.hit_voff = ctx->hit_voff.as<uint64_t>(); A.hit_nv = ctx->hit_nv.as<uint8_t>();
        A.hit_hash = mode == WALK_MODE_ALL ? ctx->hit_hash.as<uint64_t>() : nullptr;
        A.vtx_pool = ctx->vtx_pool.as<int32_t>(); A.hit_cap = hit_cap; A.vtx_cap = vtx_cap; A.ctr = d_ctr;
        A.probe = mode == WALK_MODE_PROBE ? ctx->probe.as<uint4>() : nullptr;
        // every emitted minimizer looks up one bucket of the radix directory and then the spectrum: with the directory resident in L2
        // (persisting access window) a probe costs one scattered DRAM access instead of two dependent ones
        const bool l2_pin = ctx->l2_persist_bytes > 0 && mode == WALK_MODE_PROBE;
        if (l2_pin) {
            cudaStreamAttrValue av; memset(&av, 0, sizeof av);
            av.accessPolicyWindow.base_ptr = ctx->dir.p;
            av.accessPolicyWindow.num_bytes = std::min<size_t>(((size_t)1 << dbits) * 4 + 4, ctx->l2_window_max);
            av.accessPolicyWindow.hitRatio = std::min(1.0f, (float)ctx->l2_persist_bytes / (float)std::max<size_t>(av.accessPolicyWindow.num_bytes, 1));
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            CU(cudaStreamSetAttribute(ctx->st, cudaStreamAttributeAccessPolicyWindow, &av));
        }
        CU(cudaEventRecord(ctx->ev[EV_WK0], ctx->st));
        CU(launch_walk_sketch(A, NT_, ctx->st)); ctx->launches++;
        CU(cudaEventRecord(ctx->ev[EV_WK1], ctx->st));
        if (l2_pin) {
            cudaStreamAttrValue av; memset(&av, 0, sizeof av);
            CU(cudaStreamSetAttribute(ctx->st, cudaStreamAttributeAccessPolicyWindow, &av));
        }
        CU(read_counters(ctx));
        if (l2_pin) cudaCtxResetPersistingL2Cache();
        o.n_hits = ctx->h_ctr[CTR_HITS]; o.n_hit_vtx = ctx->h_ctr[CTR_HIT_VTX];
        if (o.n_hits <= hit_cap && o.n_hit_vtx <= vtx_cap) break;
        if (attempt) return ctx->fail(PHI_ERR_CUDA, "walk hit buffers overflowed twice (internal error)");
        hit_cap = o.n_hits + 1024; vtx_cap = o.n_hit_vtx + 1024;          // exact sizes are now known: run again
    }
    // every member chunk emits what its representative emitted: per-walk minimizer counts, total hits
    CU(cudaMemsetAsync(d_ctr + CTR_PATH_HITS, 0, 8, ctx->st));
    CU(chunk_emitted(chunk_table(ctx), ctx->c_emitted.as<uint32_t>(), ctx->c_hits.as<uint32_t>(), wbase, ctx->mpw.as<unsigned long long>(), d_ctr,
                     ctx->st, &ctx->launches));
    return PHI_OK;
}

// ---- stage: threshold filter, final order, CSR
static FilterArgs filter_args(phi_gpu_index_ctx *ctx, uint64_t n, const DevBuf &rank, const DevBuf &walk, const DevBuf &pos, const DevBuf &voff,
                              const DevBuf &nv, const DevBuf &vtx, uint32_t n_spec, float threshold, uint32_t n_walks_global)
{
    FilterArgs A;
    A.n_hits = n; A.hit_rank = rank.as<uint32_t>(); A.hit_walk = walk.as<uint32_t>(); A.hit_pos = pos.as<uint32_t>();
    A.hit_voff = voff.as<uint64_t>(); A.hit_nv = nv.as<uint8_t>(); A.vtx_pool = vtx.as<int32_t>();
    A.n_ranks = n_spec;
    A.thr = threshold * (float)n_walks_global;                            // float * uint32 -> float, as ILP_index.cpp:698
    A.rank_bits = bits_for(n_spec ? n_spec - 1 : 0);
    return A;
}

// group table over the records of A; a record stands for weight[i] occurrences, or for as many as its chunk has members
// (chunk_weight), or for one.  Fills W.hit_slot / g_rep / g_cnt.
// owner_side: the table of the summaries a rank owner received (several GPUs); the local table stays alive next to it.
static int count_groups_adaptive(phi_gpu_index_ctx *ctx, const FilterArgs &A, FilterWork &W, const uint32_t *weight, const uint32_t *chunk_weight,
                                 bool owner_side)
{
    unsigned long long *d_ctr = ctx->ctr.as<unsigned long long>();
    const uint64_t n = A.n_hits;
    // distinct (rank, vertex list) groups are usually fewer than records, so start small and grow on overflow
    // (a table twice this size, load <= 50 % whatever the data, was measured on c4: the filter stage went from 11.6 to 14.4 ms —
    // the fill and the wider scatter cost more than the saved probes)
    uint64_t gcap = 1024; while (gcap < n / 2) gcap <<= 1;
    uint64_t &hint = owner_side ? ctx->gcap_hint2 : ctx->gcap_hint;
    if (owner_side) { gcap = 1024; while (gcap < 2 * n) gcap <<= 1; }     // at most n groups: the 80 % load limit is out of reach, no host wait needed
    else if (hint > gcap) gcap = hint;
    DevBuf &slotb = owner_side ? ctx->hit_slot2 : ctx->hit_slot, &tabb = owner_side ? ctx->g_slot2 : ctx->g_slot, &probeb = owner_side ? ctx->probe2 : ctx->probe;
    CU(slotb.reserve(n * 4 + 4));
    W.hit_slot = slotb.as<uint32_t>();
    W.hit_sub = nullptr;
    if (!owner_side) { CU(ctx->hit_sub.reserve(n * 4 + 4)); W.hit_sub = ctx->hit_sub.as<uint32_t>(); }
    W.weight = weight; W.chunk_weight = chunk_weight;
    if (owner_side) {                                                     // received summaries: probe records from their SoA arrays
        CU(probeb.reserve(n * 32 + 32));
        CU(filter_build_probe(A, probeb.as<uint4>(), ctx->st, &ctx->launches));
    }                                                                     // (the walk kernel wrote the probe records of its own hits)
    W.probe = probeb.as<uint4>();
    for (;;) {
        CU(tabb.reserve(gcap * 8));
        CU(fill_u64(tabb.as<uint64_t>(), gcap, 0x00000000FFFFFFFFull, ctx->st, &ctx->launches));   // (representative: none, count: 0)
        CU(cudaMemsetAsync(d_ctr + CTR_GROUPS, 0, 2 * 8, ctx->st));
        W.g_slot = tabb.as<uint2>(); W.g_cap = gcap;
        CU(filter_count_groups(A, W, ctx->st, &ctx->launches));
        if (owner_side) break;
        CU(read_counters(ctx));
        if (!ctx->h_ctr[CTR_GROUP_OVERFLOW]) break;
        gcap <<= 2;
        hint = gcap;
    }
    return PHI_OK;
}

// Instantiate the hits of the representative chunks whose rank survived for every member chunk -> ctx->x_* in
// (walk, position) order; ns = number of records.
static int expand_survivors(phi_gpu_index_ctx *ctx, int w, bool with_hash, uint64_t &ns)
{
    ns = 0;
    const uint32_t NC = ctx->n_chunks;
    if (!NC) return PHI_OK;
    ChunkTable C = chunk_table(ctx);
    CU(chunk_survivors(C, ctx->tiles.as<TileRec>(), ctx->n_tiles, ctx->hseg_off.as<uint32_t>(), ctx->hseg_cnt.as<uint32_t>(), ctx->hit_rank.as<uint32_t>(),
                       ctx->hit_nv.as<uint8_t>(), ctx->rank_drop.as<uint8_t>(), ctx->c_surv.as<uint32_t>(), ctx->c_surv_vtx.as<uint32_t>(),
                       ctx->member_cnt.as<uint32_t>(), ctx->ctr.as<unsigned long long>(), ctx->st, &ctx->launches));
    CU(cudaMemsetAsync(ctx->member_cnt.as<uint32_t>() + NC, 0, 4, ctx->st));
    CU(ctx->scan_scr.reserve(scan_u32_to_u64_scratch((uint64_t)NC + 2)));
    CU(scan_u32_to_u64(ctx->member_cnt.as<uint32_t>(), ctx->member_off.as<uint64_t>(), (uint64_t)NC + 1, ctx->scan_scr.p, ctx->st, &ctx->launches));
    CU(cudaMemcpyAsync(&ns, ctx->member_off.as<uint64_t>() + NC, 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(read_counters(ctx));                                               // also: CTR_SURV_VTX (vertices of the records), CTR_FILTERED
    if (ns >= (1ull << 32)) return ctx->fail(PHI_ERR_UNSUPPORTED, "more than 2^32-1 anchors on one GPU; shard the walks over more GPUs");
    if (!ns) return PHI_OK;
    CU(ctx->x_rank.reserve(ns * 4)); CU(ctx->x_walk.reserve(ns * 4)); CU(ctx->x_pos.reserve(ns * 4)); CU(ctx->x_voff.reserve(ns * 8)); CU(ctx->x_nv.reserve(ns));
    if (with_hash) CU(ctx->x_hash.reserve(ns * 8));
    ExpandArgs X;
    X.w = w; X.walk_id_base = ctx->world > 1 ? ctx->walk_id_base : 0;
    X.member_off = ctx->member_off.as<uint64_t>(); X.hseg_off = ctx->hseg_off.as<uint32_t>(); X.hseg_cnt = ctx->hseg_cnt.as<uint32_t>();
    X.rank_drop = ctx->rank_drop.as<uint8_t>();
    X.hit_rank = ctx->hit_rank.as<uint32_t>(); X.hit_pos = ctx->hit_pos.as<uint32_t>(); X.hit_voff = ctx->hit_voff.as<uint64_t>();
    X.hit_nv = ctx->hit_nv.as<uint8_t>(); X.hit_hash = with_hash ? ctx->hit_hash.as<uint64_t>() : nullptr;
    X.x_rank = ctx->x_rank.as<uint32_t>(); X.x_walk = ctx->x_walk.as<uint32_t>(); X.x_pos = ctx->x_pos.as<uint32_t>();
    X.x_voff = ctx->x_voff.as<uint64_t>(); X.x_nv = ctx->x_nv.as<uint8_t>(); X.x_hash = with_hash ? ctx->x_hash.as<uint64_t>() : nullptr;
    CU(chunk_expand(C, X, ctx->st, &ctx->launches));
    return PHI_OK;
}

// records of A (all of them survive) -> final (rank, walk, j) order -> CSR in ctx->anchor_*
static int order_and_csr(phi_gpu_index_ctx *ctx, const FilterArgs &A, bool key_order, bool write_rank_off, uint32_t n_walks_global, RunOut &o)
{
    unsigned long long *d_ctr = ctx->ctr.as<unsigned long long>();
    const uint64_t ns = A.n_hits;
    o.n_surv = ns;
    CU(ctx->anchor_off.reserve((ns + 1) * 8));
    if (!ns) { CU(cudaMemsetAsync(ctx->anchor_off.p, 0, 8, ctx->st)); return PHI_OK; }
    FilterWork W; memset(&W, 0, sizeof(W));
    W.rank_drop = ctx->rank_drop.as<uint8_t>(); W.ctr = d_ctr;
    CU(cudaMemsetAsync(d_ctr + CTR_BIG_GROUPS, 0, 8, ctx->st));
    CU(ctx->keys_a.reserve(ns * 8)); CU(ctx->keys_b.reserve(ns * 8)); CU(ctx->vals_a.reserve(ns * 4)); CU(ctx->vals_b.reserve(ns * 4));
    CU(ctx->sort_scr.reserve(std::max(radix_sort_scratch(ns), radix_sort_u32_scratch(ns))));
    CU(ctx->scan_scr.reserve(std::max(scan_u32_scratch(ns), scan_u32_to_u64_scratch(ns + 1))));
    W.keys_a = ctx->keys_a.as<uint64_t>(); W.keys_b = ctx->keys_b.as<uint64_t>();
    W.vals_a = ctx->vals_a.as<uint32_t>(); W.vals_b = ctx->vals_b.as<uint32_t>(); W.sort_scratch = ctx->sort_scr.p;
    CU(filter_sort_records(A, W, ctx->st, &ctx->launches));
    uint32_t *order = W.vals_a;

    if (key_order) {                                                      // (rank, walk) groups with several hits: std::map<std::string> order (:680-709)
        const uint32_t big_cap = (uint32_t)(ns / 48 + 1);
        CU(ctx->big_list.reserve((size_t)big_cap * 8)); CU(ctx->tmp_order.reserve(ns * 4));
        CU(filter_fix_multi(A, order, ns, 1, ctx->big_list.as<uint32_t>(), big_cap, d_ctr, ctx->st, &ctx->launches));
        CU(filter_fix_big(A, order, ctx->tmp_order.as<uint32_t>(), ctx->big_list.as<uint32_t>(), big_cap, d_ctr, ctx->st, &ctx->launches));
    }
    CU(ctx->nv_out.reserve((ns + 1) * 4)); CU(ctx->anchor_len.reserve(ns + 4));
    CU(cudaMemsetAsync(ctx->nv_out.as<uint32_t>() + ns, 0, 4, ctx->st));
    CU(filter_csr_sizes(A, order, ns, ctx->nv_out.as<uint32_t>(), ctx->anchor_len.as<uint8_t>(), ctx->st, &ctx->launches));
    CU(scan_u32_to_u64(ctx->nv_out.as<uint32_t>(), ctx->anchor_off.as<uint64_t>(), ns + 1, ctx->scan_scr.p, ctx->st, &ctx->launches));
    const uint64_t total_vtx = ctx->h_ctr[CTR_SURV_VTX];                  // counted with the survivors (expand_survivors): no host round trip here
    o.n_anchor_vtx = total_vtx;
    CU(ctx->anchor_walk.reserve(ns * 4)); CU(ctx->anchor_vtx.reserve(total_vtx * 4 + 4));
    CU(filter_csr_fill(A, order, ns, ctx->anchor_off.as<uint64_t>(), write_rank_off ? ctx->rank_off.as<uint64_t>() : nullptr, ctx->anchor_walk.as<int32_t>(),
                       ctx->anchor_vtx.as<int32_t>(), ctx->apw.as<unsigned long long>(), n_walks_global, ctx->st, &ctx->launches));
    return PHI_OK;
}

// The surviving groups of the local table W (rank_drop is final) -> the grouped result in ctx->rank_off (u32), anchor_len
// (= group_len), anchor_vtx (= group_vtx), grp_moff (= group_member_off), anchor_walk (= member_walk), apw.
static int groups_out(phi_gpu_index_ctx *ctx, const FilterArgs &A, FilterWork &W, uint32_t n_walks_global, RunOut &o)
{
    unsigned long long *d_ctr = ctx->ctr.as<unsigned long long>();
    const uint64_t n = A.n_hits;
    CU(cudaMemsetAsync(d_ctr + CTR_SURVIVORS, 0, 2 * 8, ctx->st));       // SURVIVORS, BIG_GROUPS
    CU(cudaMemsetAsync(d_ctr + CTR_SURV_VTX, 0, 2 * 8, ctx->st));        // SURV_VTX, OUT_GROUPS
    CU(ctx->flags.reserve(n * 4 + 4)); CU(ctx->vals_b.reserve(n * 4 + 4));
    CU(ctx->keys_a.reserve(n * 4 + 4)); CU(ctx->vals_a.reserve(n * 4 + 4));
    CU(ctx->scan_scr.reserve(std::max(scan_u32_scratch(n + 2), (size_t)1024)));
    // upper bound n groups: keys_a / vals_a are sized for it, the second sort buffers after the count is known
    CU(groups_compact(A, W, ctx->flags.as<uint32_t>(), ctx->vals_b.as<uint32_t>(), ctx->keys_a.as<uint32_t>(), ctx->vals_a.as<uint32_t>(),
                      ctx->scan_scr.p, ctx->st, &ctx->launches));
    CU(read_counters(ctx));                                               // groups, members, vertices (also: CTR_FILTERED)
    const uint64_t ng = ctx->h_ctr[CTR_OUT_GROUPS], ns = ctx->h_ctr[CTR_SURVIVORS], nv = ctx->h_ctr[CTR_SURV_VTX];
    o.n_filtered = (int64_t)ctx->h_ctr[CTR_FILTERED];
    if (ns >= (1ull << 32) || nv >= (1ull << 32)) return ctx->fail(PHI_ERR_UNSUPPORTED, "more than 2^32-1 anchors on one GPU; shard the walks over more GPUs");
    o.n_surv = ns; o.n_groups = ng; o.n_anchor_vtx = nv;
    if (!ng) return PHI_OK;
    CU(ctx->keys_b.reserve(ng * 4 + 4)); CU(ctx->vals_b.reserve(ng * 4 + 4));
    CU(ctx->sort_scr.reserve(radix_sort_u32_scratch(ng)));
    CU(radix_sort_u32(ctx->keys_a.as<uint32_t>(), ctx->keys_b.as<uint32_t>(), ctx->vals_a.as<uint32_t>(), ctx->vals_b.as<uint32_t>(), ng, A.rank_bits,
                      ctx->sort_scr.p, ctx->st, &ctx->launches));
    uint32_t *order = ctx->vals_a.as<uint32_t>();
    // the groups of one rank: std::map<std::string> order of their keys (:680-709)
    const uint32_t big_cap = (uint32_t)(ng / 48 + 1);
    CU(ctx->big_list.reserve((size_t)big_cap * 8)); CU(ctx->tmp_order.reserve(ng * 4));
    CU(filter_fix_multi(A, order, ng, 0, ctx->big_list.as<uint32_t>(), big_cap, d_ctr, ctx->st, &ctx->launches));
    CU(filter_fix_big(A, order, ctx->tmp_order.as<uint32_t>(), ctx->big_list.as<uint32_t>(), big_cap, d_ctr, ctx->st, &ctx->launches));
    // sizes -> offsets
    CU(ctx->grp_cnt.reserve((ng + 1) * 4)); CU(ctx->nv_out.reserve((ng + 1) * 4)); CU(ctx->anchor_len.reserve(ng + 4));
    CU(ctx->grp_moff.reserve((ng + 1) * 4)); CU(ctx->grp_voff.reserve((ng + 1) * 4));
    CU(cudaMemsetAsync(ctx->grp_cnt.as<uint32_t>() + ng, 0, 4, ctx->st)); CU(cudaMemsetAsync(ctx->nv_out.as<uint32_t>() + ng, 0, 4, ctx->st));
    CU(groups_sizes(A, W, order, (uint32_t)ng, ctx->grp_cnt.as<uint32_t>(), ctx->nv_out.as<uint32_t>(), ctx->anchor_len.as<uint8_t>(),
                    ctx->rank_off.as<uint32_t>(), ctx->st, &ctx->launches));
    CU(ctx->scan_scr.reserve(scan_u32_scratch(ng + 2)));
    CU(scan_u32(ctx->grp_cnt.as<uint32_t>(), ctx->grp_moff.as<uint32_t>(), ng + 1, ctx->scan_scr.p, ctx->st, &ctx->launches));
    CU(scan_u32(ctx->nv_out.as<uint32_t>(), ctx->grp_voff.as<uint32_t>(), ng + 1, ctx->scan_scr.p, ctx->st, &ctx->launches));
    CU(ctx->members_tmp.reserve(ns * 4 + 4)); CU(ctx->anchor_walk.reserve(ns * 4 + 4)); CU(ctx->anchor_vtx.reserve(nv * 4 + 4));
    GroupOut G;
    G.order = order; G.member_off = ctx->grp_moff.as<uint32_t>(); G.vtx_off = ctx->grp_voff.as<uint32_t>();
    G.cm_off = ctx->cm_off.as<uint32_t>(); G.cm_walk = ctx->cm_walk.as<uint32_t>(); G.members_tmp = ctx->members_tmp.as<uint32_t>();
    G.member_walk = ctx->anchor_walk.p; G.member_walk_bytes = o.member_walk_bytes;
    G.group_vtx = ctx->anchor_vtx.as<int32_t>();
    G.anchors_per_walk = ctx->apw.as<unsigned long long>(); G.walk_id_base = ctx->world > 1 ? ctx->walk_id_base : 0;
    CU(groups_fill(A, W, G, (uint32_t)ng, ctx->n_walks, n_walks_global, ctx->st, &ctx->launches));
    return PHI_OK;
}

__global__ void summary_flags_kernel(const uint2 *g_slot, const uint32_t *hit_slot, uint64_t n, uint32_t *flags)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) flags[i] = g_slot[hit_slot[i]].x == (uint32_t)i ? 1u : 0u;
}
// one summary per local group: (rank, count, vertex list of the representative)
__global__ void summary_emit_kernel(FilterArgs A, const uint2 *g_slot, const uint32_t *hit_slot, const uint64_t *pos,
                                    uint32_t *m_rank, uint32_t *m_cnt, uint64_t *m_voff, uint8_t *m_nv)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= A.n_hits) return;
    uint32_t slot = hit_slot[i];
    const uint2 gs = g_slot[slot];
    if (gs.x != (uint32_t)i) return;
    uint64_t j = pos[i];
    m_rank[j] = A.hit_rank[i]; m_cnt[j] = gs.y; m_voff[j] = A.hit_voff[i]; m_nv[j] = A.hit_nv[i];
}

// status (several GPUs only): error code of the walk stage of this rank; it travels with the small collective of the record
// exchange so that all ranks stop together.
static int stage_filter(phi_gpu_index_ctx *ctx, int w, int mode, uint32_t n_walks_global, float threshold, RunOut &o, int status)
{
    if (status && ctx->world > 1 && mode == WALK_MODE_PROBE) {
        RouteIn none; memset(&none, 0, sizeof none);
        uint64_t a = 0, b = 0;
        int rc = exchange_records(ctx, none, a, b, status);
        return rc ? rc : status;
    }
    unsigned long long *d_ctr = ctx->ctr.as<unsigned long long>();
    CU(ctx->apw.reserve(((size_t)n_walks_global + 1) * 8));
    CU(cudaMemsetAsync(ctx->apw.p, 0, ((size_t)n_walks_global + 1) * 8, ctx->st));
    CU(cudaMemsetAsync(d_ctr + CTR_FILTERED, 0, 3 * 8, ctx->st));        // FILTERED, SURVIVORS, BIG_GROUPS
    CU(ctx->rank_drop.reserve((size_t)o.n_spec + 4));
    CU(cudaMemsetAsync(ctx->rank_drop.p, 0, (size_t)o.n_spec + 4, ctx->st));
    CU(ctx->anchor_off.reserve(8));
    CU(cudaMemsetAsync(ctx->anchor_off.p, 0, 8, ctx->st));
    CU(ctx->rank_off.reserve(((size_t)o.n_spec + 2) * 8));
    CU(cudaMemsetAsync(ctx->rank_off.p, 0, ((size_t)o.n_spec + 1) * 8, ctx->st));   // no anchors: every rank is empty
    o.n_surv = 0; o.n_anchor_vtx = 0; o.n_filtered = 0;
    o.member_walk_bytes = (mode == WALK_MODE_PROBE && n_walks_global <= 65536) ? 2 : 4;
    FilterWork W; memset(&W, 0, sizeof(W));
    W.rank_drop = ctx->rank_drop.as<uint8_t>(); W.ctr = d_ctr;
    const uint64_t n = o.n_hits;                                          // hits of the representative chunks
    // the representatives' hits: hit_chunk takes the place of the walk id (its member count is the record's weight)
    FilterArgs A = filter_args(ctx, n, ctx->hit_rank, ctx->hit_chunk, ctx->hit_pos, ctx->hit_voff, ctx->hit_nv, ctx->vtx_pool, o.n_spec, threshold,
                               n_walks_global);

    if (mode == WALK_MODE_ALL) {
        // sketch-only: every emitted minimizer of every walk, in (walk, position) order, no filter
        uint64_t ns = 0;
        int rc = expand_survivors(ctx, w, true, ns);
        if (rc) return rc;
        FilterArgs X = filter_args(ctx, ns, ctx->x_rank, ctx->x_walk, ctx->x_pos, ctx->x_voff, ctx->x_nv, ctx->vtx_pool, 1, 0.f, n_walks_global);
        X.rank_bits = 0;                                                  // already in final order
        return order_and_csr(ctx, X, false, false, n_walks_global, o);
    }

    // ---- which ranks are dropped.  One GPU: local group counts are the global ones.  Several GPUs: one (rank, count, list)
    // summary per local group travels to the owner of the rank, which adds the partial counts up and applies the threshold;
    // the drop flags of all owners are then shared.  Every rank takes part in every collective, also with zero hits.
    if (ctx->world == 1) {
        if (!n) return PHI_OK;
        W.mark_inline = 1;                                                // one GPU: the local counts are the global ones
        int rc = count_groups_adaptive(ctx, A, W, nullptr, ctx->c_ninst.as<uint32_t>(), false);
        if (rc) return rc;
    } else {
        std::string err; NcclApi *nc = nccl_api(err);
        if (!nc || !ctx->comm) return ctx->fail(PHI_ERR_COMM, "communicator not initialised");
        CU(cudaEventRecord(ctx->ev[EV_XH0], ctx->st));
        uint64_t n_sum = 0;
        auto summaries = [&]() -> int {                                       // local group table -> one summary per local group
            if (!n) return PHI_OK;
            int rc = count_groups_adaptive(ctx, A, W, nullptr, ctx->c_ninst.as<uint32_t>(), false);
            if (rc) return rc;
            CU(ctx->flags.reserve(n * 4 + 4)); CU(ctx->flags64.reserve((n + 1) * 8));
            CU(ctx->scan_scr.reserve(std::max(scan_u32_scratch(n), scan_u32_to_u64_scratch(n + 1))));
            summary_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->st>>>(W.g_slot, W.hit_slot, n, ctx->flags.as<uint32_t>());
            CU(cudaGetLastError()); ctx->launches++;
            CU(scan_u32_to_u64(ctx->flags.as<uint32_t>(), ctx->flags64.as<uint64_t>(), n, ctx->scan_scr.p, ctx->st, &ctx->launches));
            n_sum = ctx->h_ctr[CTR_GROUPS];                                   // distinct local groups (read by count_groups_adaptive)
            CU(ctx->m_rank.reserve(n_sum * 4 + 4)); CU(ctx->m_cnt.reserve(n_sum * 4 + 4)); CU(ctx->m_voff.reserve(n_sum * 8 + 8)); CU(ctx->m_nv.reserve(n_sum + 4));
            summary_emit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->st>>>(A, W.g_slot, W.hit_slot, ctx->flags64.as<uint64_t>(),
                                                                                 ctx->m_rank.as<uint32_t>(), ctx->m_cnt.as<uint32_t>(), ctx->m_voff.as<uint64_t>(), ctx->m_nv.as<uint8_t>());
            CU(cudaGetLastError()); ctx->launches++;
            return PHI_OK;
        };
        const int rc_local = summaries();
        RouteIn I;
        I.rank = ctx->m_rank.as<uint32_t>(); I.cnt = ctx->m_cnt.as<uint32_t>(); I.voff = ctx->m_voff.as<uint64_t>();
        I.nv = ctx->m_nv.as<uint8_t>(); I.vtx = ctx->vtx_pool.as<int32_t>(); I.n = rc_local ? 0 : n_sum;
        CU(cudaEventRecord(ctx->ev[EV_XH1], ctx->st));
        uint64_t rs = 0, rsv = 0;
        int rc = exchange_records(ctx, I, rs, rsv, rc_local);
        if (rc) return rc;
        if (rs) {                                                             // owner: add the partial counts up, apply the threshold
            FilterArgs B = filter_args(ctx, rs, ctx->r_rank, ctx->r_walk, ctx->r_walk, ctx->r_voff, ctx->r_nv, ctx->r_vtx, o.n_spec, threshold, n_walks_global);
            FilterWork WB; memset(&WB, 0, sizeof(WB));
            WB.rank_drop = ctx->rank_drop.as<uint8_t>(); WB.ctr = d_ctr;
            WB.mark_inline = 1;                                                // drops (and counts) the ranks of this owner's range
            rc = count_groups_adaptive(ctx, B, WB, ctx->r_walk.as<uint32_t>(), nullptr, true);
            if (rc) { comm_release(ctx, true); return rc; }
        }
        // share the drop flags: every owner marked ranks of its own range only, so the byte-wise maximum over the GPUs is the truth
        if (o.n_spec) NC(nc->AllReduce(ctx->rank_drop.p, ctx->rank_drop.p, (size_t)o.n_spec, ncclUint8, ncclMax, (ncclComm_t)ctx->comm, ctx->st));
        CU(cudaEventRecord(ctx->ev[EV_XH2], ctx->st));
        if (!n) { CU(read_counters(ctx)); o.n_filtered = (int64_t)ctx->h_ctr[CTR_FILTERED]; return PHI_OK; }
    }

    // ---- the surviving groups of THIS GPU's walks, in (rank, key) order, with their member walks
    return groups_out(ctx, A, W, n_walks_global, o);
}

// ---- the -d1 statistic (ILP_index.cpp:565-606), after the result proper is complete: every minimizer of the representative
// chunks with its hash -> instantiated per walk in (walk, position) order -> stable sort on the hash (walks stay ascending
// inside a hash) -> distinct walks per hash -> histogram.  Reuses the hit / expansion / sort buffers.
__global__ void gather_u64_kernel(const uint64_t *src, const uint32_t *idx, uint64_t n, uint64_t *dst)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void iota_u32_kernel(uint32_t *p, uint64_t n)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

// Several GPUs: the (hash, walk) pairs of this GPU's walk slices (ns of them, sorted by hash, in x_hash / x_walk) travel to the owner
// of the hash (the range partition of the spectrum exchange); the owner orders what it received by (hash, walk).  ns becomes the
// number of pairs this GPU owns.  The same pair can arrive from two GPUs (one walk carrying a minimizer in two regions): the
// statistic counts distinct walks per hash, so duplicates do not matter once the walks of a hash are sorted.
static int exchange_hash_walk_pairs(phi_gpu_index_ctx *ctx, uint64_t &ns)
{
    std::string err; NcclApi *nc = nccl_api(err);
    if (!nc || !ctx->comm) return ctx->fail(PHI_ERR_COMM, "communicator not initialised");
    const int W = ctx->world, me = ctx->rank;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    CU(ctx->xcnt.reserve(8192));
    uint64_t *d_split = ctx->xcnt.as<uint64_t>();
    owner_split_kernel<<<1, 128, 0, ctx->st>>>(ctx->x_hash.as<uint64_t>(), ns, W, 0, d_split);
    CU(cudaGetLastError()); ctx->launches++;
    const uint64_t *all = nullptr;
    int rc = allgather_words(ctx, nc, d_split, W + 2, all);
    if (rc) return rc;
    std::vector<uint64_t> scnt(W), soff(W), rcnt(W), roff(W);
    uint64_t rtot = 0;
    for (int p = 0; p < W; ++p) {
        const uint64_t *sp = all + (size_t)p * (W + 2), *mine = all + (size_t)me * (W + 2);
        scnt[p] = mine[p + 1] - mine[p]; soff[p] = mine[p];
        rcnt[p] = sp[me + 1] - sp[me]; roff[p] = rtot; rtot += rcnt[p];
    }
    if (rtot >= (1ull << 32)) return ctx->fail(PHI_ERR_UNSUPPORTED, "more than 2^32-1 (hash, walk) pairs routed to one GPU");
    struct AbortGuard { phi_gpu_index_ctx *c; bool armed; ~AbortGuard() { if (armed) comm_release(c, true); } } guard = {ctx, true};
    CU(ctx->keys_a.reserve(rtot * 8 + 8)); CU(ctx->vals_a.reserve(rtot * 4 + 4));
    NC(nc->GroupStart());
    for (int p = 0; p < W; ++p) {
        if (p == me) continue;
        if (scnt[p]) { NC(nc->Send(ctx->x_hash.as<uint64_t>() + soff[p], scnt[p], ncclUint64, p, comm, ctx->st)); NC(nc->Send(ctx->x_walk.as<uint32_t>() + soff[p], scnt[p], ncclUint32, p, comm, ctx->st)); }
        if (rcnt[p]) { NC(nc->Recv(ctx->keys_a.as<uint64_t>() + roff[p], rcnt[p], ncclUint64, p, comm, ctx->st)); NC(nc->Recv(ctx->vals_a.as<uint32_t>() + roff[p], rcnt[p], ncclUint32, p, comm, ctx->st)); }
    }
    NC(nc->GroupEnd());
    if (scnt[me]) {
        CU(cudaMemcpyAsync(ctx->keys_a.as<uint64_t>() + roff[me], ctx->x_hash.as<uint64_t>() + soff[me], scnt[me] * 8, cudaMemcpyDeviceToDevice, ctx->st));
        CU(cudaMemcpyAsync(ctx->vals_a.as<uint32_t>() + roff[me], ctx->x_walk.as<uint32_t>() + soff[me], scnt[me] * 4, cudaMemcpyDeviceToDevice, ctx->st));
    }
    guard.armed = false;
    ns = rtot;
    if (!ns) return PHI_OK;
    // (hash, walk) order: stable sort on the walk first (carrying the pair index), then on the hash
    CU(ctx->x_hash.reserve(ns * 8)); CU(ctx->x_walk.reserve(ns * 4)); CU(ctx->keys_b.reserve(ns * 8)); CU(ctx->vals_b.reserve(ns * 4)); CU(ctx->tmp_order.reserve(ns * 4));
    CU(ctx->sort_scr.reserve(std::max(radix_sort_scratch(ns), radix_sort_u32_scratch(ns))));
    const unsigned nb = (unsigned)((ns + 255) / 256);
    iota_u32_kernel<<<nb, 256, 0, ctx->st>>>(ctx->tmp_order.as<uint32_t>(), ns);
    CU(cudaGetLastError()); ctx->launches++;
    CU(radix_sort_u32(ctx->vals_a.as<uint32_t>(), ctx->vals_b.as<uint32_t>(), ctx->tmp_order.as<uint32_t>(), ctx->x_walk.as<uint32_t>(), ns, 32, ctx->sort_scr.p, ctx->st, &ctx->launches));
    gather_u64_kernel<<<nb, 256, 0, ctx->st>>>(ctx->keys_a.as<uint64_t>(), ctx->tmp_order.as<uint32_t>(), ns, ctx->x_hash.as<uint64_t>());
    CU(cudaGetLastError()); ctx->launches++;
    CU(cudaMemcpyAsync(ctx->x_walk.p, ctx->vals_a.p, ns * 4, cudaMemcpyDeviceToDevice, ctx->st));       // the walks, ascending
    CU(radix_sort_u64(ctx->x_hash.as<uint64_t>(), ctx->keys_b.as<uint64_t>(), ctx->x_walk.as<uint32_t>(), ctx->vals_b.as<uint32_t>(), ns, 0, 64,
                      ctx->sort_scr.p, ctx->st, &ctx->launches));
    return PHI_OK;
}

static int stage_debug_hist(phi_gpu_index_ctx *ctx, int k, int w, const std::vector<uint64_t> &h_walk_len, const uint32_t *d_walk_vtx,
                            const uint64_t *d_walk_off, uint64_t n_steps_eff, int walks_monotone, uint32_t n_walks_global)
{
    unsigned long long *d_ctr = ctx->ctr.as<unsigned long long>();
    CU(ctx->dbg_hist.reserve(((size_t)n_walks_global + 2) * 8));
    CU(cudaMemsetAsync(ctx->dbg_hist.p, 0, ((size_t)n_walks_global + 2) * 8, ctx->st));
    CU(cudaMemsetAsync(d_ctr + CTR_WALK_KMERS, 0, 8, ctx->st));
    RunOut tmp;
    int rc = stage_walks(ctx, k, w, WALK_MODE_ALL, 0, h_walk_len, d_walk_vtx, d_walk_off, n_steps_eff, walks_monotone, tmp);
    if (rc) return rc;
    CU(ctx->rank_drop.reserve(4)); CU(cudaMemsetAsync(ctx->rank_drop.p, 0, 4, ctx->st));   // every record has rank 0 here: keep it
    uint64_t ns = 0;
    rc = expand_survivors(ctx, w, true, ns);
    if (rc) return rc;                                                     // (several GPUs: every rank reaches the collectives below or none does — the
    if (ns) {                                                              //  failures above are allocation failures, which abort the run on this rank)
        CU(ctx->keys_b.reserve(ns * 8)); CU(ctx->vals_b.reserve(ns * 4)); CU(ctx->sort_scr.reserve(radix_sort_scratch(ns)));
        CU(radix_sort_u64(ctx->x_hash.as<uint64_t>(), ctx->keys_b.as<uint64_t>(), ctx->x_walk.as<uint32_t>(), ctx->vals_b.as<uint32_t>(), ns, 0, 64,
                          ctx->sort_scr.p, ctx->st, &ctx->launches));
    }
    unsigned long long *d_hist = ctx->dbg_hist.as<unsigned long long>();
    if (ctx->world > 1) {
        if (!ns) { CU(ctx->x_hash.reserve(8)); CU(ctx->x_walk.reserve(8)); }
        rc = exchange_hash_walk_pairs(ctx, ns);
        if (rc) return rc;
    }
    if (ns) CU(filter_shared_kmer_hist(ctx->x_hash.as<uint64_t>(), ctx->x_walk.as<uint32_t>(), ns, n_walks_global, d_hist, d_hist + n_walks_global + 1,
                                       ctx->st, &ctx->launches));
    if (ctx->world > 1) {                                                  // every hash has one owner: the histograms and the distinct counts add up
        std::string err; NcclApi *nc = nccl_api(err);
        NC(nc->AllReduce(d_hist, d_hist, (size_t)n_walks_global + 2, ncclUint64, ncclSum, (ncclComm_t)ctx->comm, ctx->st));
    }
    CU(cudaMemcpyAsync(d_ctr + CTR_WALK_KMERS, d_hist + n_walks_global + 1, 8, cudaMemcpyDeviceToDevice, ctx->st));
    return PHI_OK;
}

static PinnedBuf pinned_acquire(phi_gpu_index_ctx *ctx, size_t bytes)
{
    PinnedBuf best; int bi = -1;
    for (size_t i = 0; i < ctx->pinned_pool.size(); ++i)
        if (ctx->pinned_pool[i].cap >= bytes && (bi < 0 || ctx->pinned_pool[i].cap < best.cap)) { best = ctx->pinned_pool[i]; bi = (int)i; }
    if (bi >= 0) { ctx->pinned_pool.erase(ctx->pinned_pool.begin() + bi); return best; }
    PinnedBuf b; size_t want = bytes + bytes / 4 + 4096;
    if (cudaHostAlloc(&b.p, want, cudaHostAllocDefault) == cudaSuccess) b.cap = want; else b.p = nullptr;
    return b;
}

static phi_index_result *alloc_result(phi_gpu_index_ctx *ctx)
{
    ResultBox *b = (ResultBox *)calloc(1, sizeof(ResultBox));
    if (b) b->owner = ctx;
    return b ? &b->pub : nullptr;
}

extern "C" void phi_gpu_index_result_free(phi_index_result *r)
{
    if (!r) return;
    ResultBox *b = (ResultBox *)r;
    if (b->heap) {                                                          // phi_index_result_merge: plain malloc'ed arrays
        for (int i = 0; i < b->nbufs; ++i) free(b->bufs[i].p);
        free(b);
        return;
    }
    std::lock_guard<std::mutex> lk(g_live_mu);
    const bool alive = g_live_ctx.count(b->owner) != 0;
    for (int i = 0; i < b->nbufs; ++i) {
        if (!b->bufs[i].p) continue;
        if (alive) b->owner->pinned_pool.push_back(b->bufs[i]); else cudaFreeHost(b->bufs[i].p);
    }
    free(b);
}

extern "C" void *phi_gpu_host_alloc(size_t bytes)
{
    void *p = nullptr;
    return cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
extern "C" void phi_gpu_host_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" void phi_gpu_index_free_u64(uint64_t *p) { free(p); }

static int validate_params(phi_gpu_index_ctx *ctx, const phi_index_params *p)
{
    if (!p) return ctx->fail(PHI_ERR_ARG, "params is NULL");
    if (p->k < 1 || p->w < 1) return ctx->fail(PHI_ERR_ARG, "k and w must be >= 1");
    if (p->k > 255) return ctx->fail(PHI_ERR_UNSUPPORTED, "k > 255 is not implemented on the GPU path (vertex lists carry a one-byte length); refusing rather than diverging");
    if (p->w > 256) return ctx->fail(PHI_ERR_UNSUPPORTED, "w > 256 is not implemented on the GPU path");
    return PHI_OK;
}

template <class T>
static int download(phi_gpu_index_ctx *ctx, phi_index_result *res, const void *dev, uint64_t n, const T **out, cudaStream_t stream = nullptr)
{
    if (!stream) stream = ctx->st;
    ResultBox *b = (ResultBox *)res;
    PinnedBuf pb = pinned_acquire(ctx, std::max<uint64_t>(n, 1) * sizeof(T));
    if (!pb.p) return ctx->fail(PHI_ERR_NOMEM, "pinned host allocation failed");
    b->bufs[b->nbufs++] = pb;
    if (n) CU(cudaMemcpyAsync(pb.p, dev, n * sizeof(T), cudaMemcpyDeviceToHost, stream));
    *out = (const T *)pb.p;
    return PHI_OK;
}

static void collect_times(phi_gpu_index_ctx *ctx)
{
    auto el = [&](int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]); return ms; };
    phi_stage_times &t = ctx->times;
    t.h2d_ms = el(EV_START, EV_H2D);                                     // host -> device copies overlap the read stage (two streams)
    t.graph_prep_ms = el(EV_PREP0, EV_PREP);                             // second stream, concurrent with the read stage
    t.read_sketch_ms = el(EV_RD0, EV_READS);
    t.spectrum_ms = el(EV_READS, EV_SPECTRUM);
    t.walk_sketch_ms = el(EV_SPECTRUM, EV_WALKS);
    t.filter_ms = el(EV_WALKS, EV_FILTER);
    t.d2h_ms = el(EV_FILTER, EV_END);
    t.total_ms = el(EV_START, EV_END) ;
    t.walk_kernel_ms = el(EV_WK0, EV_WK1);
    t.read_kernel_ms = el(EV_RK0, EV_RK1);
    t.kernel_launches = ctx->launches;
    t.exchange_spectrum_ms = t.route_hits_ms = t.exchange_hits_ms = 0.f;
    if (ctx->world > 1) {
        t.exchange_spectrum_ms = el(EV_XS0, EV_XS1);
        t.route_hits_ms = el(EV_XH0, EV_XH1);
        t.exchange_hits_ms = el(EV_XH1, EV_XH2);
    }
}

static int run_pipeline(phi_gpu_index_ctx *ctx, const phi_index_params *prm, int mode, int do_download, phi_index_result **out,
                        uint64_t **hashes_out, bool uploading)
{
    if (!ctx->have_inputs) return ctx->fail(PHI_ERR_ARG, "no inputs uploaded");
    int rc = validate_params(ctx, prm);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    const int k = prm->k, w = prm->w;
    ctx->launches = 0;
    if (!uploading) { CU(cudaEventRecord(ctx->ev[EV_START], ctx->st)); CU(cudaEventRecord(ctx->ev[EV_H2D], ctx->st)); }
    CU(cudaMemsetAsync(ctx->ctr.p, 0, CTR_COUNT * 8, ctx->st));
    memset(ctx->h_ctr, 0, CTR_COUNT * 8);                                 // the host mirror starts from the same state
    RunOut o;
    std::vector<uint64_t> h_walk_len; uint64_t n_steps_eff = 0;
    const uint32_t *d_walk_vtx; const uint64_t *d_walk_off;
    int walks_monotone = 1;
    // main stream: read sketch kernel in flight ...
    int dbits = 0;
    ReadsState rs;
    int status = PHI_OK;          // several GPUs: a stage that fails on this rank is reported through the next exchange (all ranks stop together)
    if (mode == WALK_MODE_PROBE) {
        rc = stage_reads_begin(ctx, k, w, rs);
        if (rc) { if (ctx->world == 1) return rc; status = rc; }
    } else {
        CU(cudaEventRecord(ctx->ev[EV_RD0], ctx->st));
        CU(cudaEventRecord(ctx->ev[EV_RK0], ctx->st)); CU(cudaEventRecord(ctx->ev[EV_RK1], ctx->st));
        CU(cudaEventRecord(ctx->ev[EV_READS], ctx->st));
    }
    // ... while a second host thread prepares the graph on the second stream: its host-side waits (number of chunks, number of
    // tiles) and the main thread's (spectrum size, and with several GPUs the whole spectrum exchange) no longer queue up behind each other
    int rc_prep = PHI_OK;
    ctx->launches2 = 0;
    std::thread prep([&]() {
        cudaSetDevice(ctx->device);
        rc_prep = stage_graph_prep(ctx, k, w, h_walk_len, d_walk_vtx, d_walk_off, n_steps_eff, walks_monotone);
        if (!rc_prep && cudaEventRecord(ctx->ev[EV_PREP], ctx->st2) != cudaSuccess) rc_prep = ctx->fail2(PHI_ERR_CUDA, "cudaEventRecord failed");
    });
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner = {prep};   // every return path waits for the thread
    if (mode == WALK_MODE_PROBE) {
        rc = stage_reads_finish(ctx, k, w, rs, o, dbits, status);
        if (rc) return rc;
    }
    prep.join();
    ctx->launches += ctx->launches2;
    if (rc_prep) {                                                         // (several GPUs: reported through the record exchange, all ranks stop together)
        ctx->err = ctx->err2;
        if (ctx->world == 1 || mode != WALK_MODE_PROBE) return rc_prep;
        status = rc_prep;
    }
    // the result starts to travel as soon as its parts exist: the spectrum goes out on the copy stream under the walk stage
    struct ResGuard { phi_index_result *r; ~ResGuard() { if (r) phi_gpu_index_result_free(r); } } guard = {alloc_result(ctx)};
    phi_index_result *res = guard.r;
    if (!res) return ctx->fail(PHI_ERR_NOMEM, "host allocation failed");
    // several GPUs: the ranked spectrum is the same on every rank; rank 0 alone copies it out (the others return spectrum == NULL)
    const bool spec_early = do_download && mode == WALK_MODE_PROBE && (ctx->world == 1 || ctx->rank == 0);
    if (spec_early) {
        CU(cudaEventRecord(ctx->ev_sync, ctx->st));
        CU(cudaStreamWaitEvent(ctx->st_copy, ctx->ev_sync, 0));
        rc = download<uint64_t>(ctx, res, ctx->spec_a.p, o.n_spec, &res->spectrum, ctx->st_copy);
        if (rc) return rc;
        CU(cudaEventRecord(ctx->ev_graph_in, ctx->st_copy));               // (the graph-arrived event is free again: reuse it as "spectrum out")
    }
    CU(cudaEventRecord(ctx->ev[EV_SPECTRUM], ctx->st));
    CU(cudaStreamWaitEvent(ctx->st, ctx->ev[EV_PREP], 0));               // the walk stage needs both
    if (!status) {
        rc = stage_walks(ctx, k, w, mode, dbits, h_walk_len, d_walk_vtx, d_walk_off, n_steps_eff, walks_monotone, o);
        if (rc) { if (ctx->world == 1 || mode != WALK_MODE_PROBE) return rc; status = rc; }
    }
    CU(cudaEventRecord(ctx->ev[EV_WALKS], ctx->st));

    const uint32_t H = ctx->n_walks;
    const uint32_t HG = ctx->world > 1 ? ctx->n_walks_global : H;

    rc = stage_filter(ctx, w, mode, HG, prm->threshold, o, status);
    if (rc) return rc;
    o.path_hits = ctx->h_ctr[CTR_PATH_HITS];                               // read back by the syncs of the filter stage
    ctx->unique_hits = o.n_hits;
    CU(cudaEventRecord(ctx->ev[EV_FILTER], ctx->st));
    const bool want_hist = mode == WALK_MODE_PROBE && prm->debug != 0;
    if (want_hist) {
        rc = stage_debug_hist(ctx, k, w, h_walk_len, d_walk_vtx, d_walk_off, n_steps_eff, walks_monotone, HG);
        if (rc) return rc;
    }

    res->count_sp_r = (int32_t)o.n_spec; res->n_walks = HG; res->n_filtered = o.n_filtered;
    if (mode == WALK_MODE_ALL) o.n_groups = o.n_surv;                       // sketch-only: one group per emitted minimizer
    res->n_anchors = o.n_surv; res->n_groups = o.n_groups; res->n_group_vtx = o.n_anchor_vtx;
    res->read_kmer_positions = o.read_pos; res->path_kmer_positions = o.path_pos;
    res->read_minimizers_emitted = o.read_emitted; res->path_hits = o.path_hits;
    if (do_download) {
        rc = (spec_early || ctx->world > 1) ? PHI_OK : download<uint64_t>(ctx, res, ctx->spec_a.p, 0, &res->spectrum);
        if (!rc && mode == WALK_MODE_PROBE) rc = download<uint32_t>(ctx, res, ctx->rank_off.p, (uint64_t)o.n_spec + 1, &res->rank_off);
        if (!rc && mode == WALK_MODE_PROBE) {
            if (o.n_groups) rc = download<uint32_t>(ctx, res, ctx->grp_moff.p, o.n_groups + 1, &res->group_member_off);
            else { rc = download<uint32_t>(ctx, res, nullptr, 0, &res->group_member_off); if (!rc) *(uint32_t *)res->group_member_off = 0; }
        }
        if (!rc && o.member_walk_bytes == 2 && mode == WALK_MODE_PROBE) rc = download<uint16_t>(ctx, res, ctx->anchor_walk.p, o.n_surv, &res->member_walk16);
        else if (!rc) rc = download<int32_t>(ctx, res, ctx->anchor_walk.p, o.n_surv, &res->member_walk32);
        if (!rc) rc = download<uint8_t>(ctx, res, ctx->anchor_len.p, o.n_groups, &res->group_len);
        if (!rc) rc = download<int32_t>(ctx, res, ctx->anchor_vtx.p, o.n_anchor_vtx, &res->group_vtx);
        if (!rc) rc = download<uint64_t>(ctx, res, ctx->apw.as<uint64_t>(), HG, &res->anchors_per_walk);
        if (rc) return rc;
    }
    if (want_hist) {
        int rc3 = download<uint64_t>(ctx, res, ctx->dbg_hist.p, (uint64_t)HG + 1, &res->shared_kmer_hist);
        if (!rc3) rc3 = read_counters(ctx) == cudaSuccess ? PHI_OK : ctx->fail(PHI_ERR_CUDA, "counter read failed");
        if (rc3) return rc3;
        res->n_walk_kmers = ctx->h_ctr[CTR_WALK_KMERS];
    }
    {   // per-walk minimizer counts are tiny and always returned
        int rc2 = download<uint64_t>(ctx, res, ctx->mpw.p, HG, &res->minimizers_per_walk);
        if (rc2) return rc2;
    }
    uint64_t *hashes = nullptr; uint32_t *h_order = nullptr;
    if (mode == WALK_MODE_ALL && hashes_out && o.n_surv) {
        hashes = (uint64_t *)malloc(o.n_surv * 8); h_order = (uint32_t *)malloc(o.n_surv * 4);
        CU(cudaMemcpyAsync(hashes, ctx->x_hash.p, o.n_surv * 8, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(h_order, ctx->vals_a.p, o.n_surv * 4, cudaMemcpyDeviceToHost, ctx->st));
    }
    if (spec_early) CU(cudaStreamWaitEvent(ctx->st, ctx->ev_graph_in, 0));   // the spectrum copy on the copy stream
    CU(cudaEventRecord(ctx->ev[EV_END], ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    for (uint32_t h = 0; h < HG; ++h) res->path_minimizers_emitted += res->minimizers_per_walk[h];
    if (mode == WALK_MODE_ALL && hashes_out) {
        uint64_t *sorted = (uint64_t *)malloc(std::max<uint64_t>(o.n_surv, 1) * 8);
        for (uint64_t i = 0; i < o.n_surv; ++i) sorted[i] = hashes[h_order[i]];
        free(hashes); free(h_order);
        *hashes_out = sorted;
    }
    collect_times(ctx);
    guard.r = nullptr;                                                      // success: the caller owns the result now
    *out = res;
    return PHI_OK;
}

extern "C" int phi_gpu_index_run_resident(phi_gpu_index_ctx *ctx, const phi_index_params *params, int download_result, phi_index_result **out)
{
    if (!ctx || !out) return PHI_ERR_ARG;
    *out = nullptr;
    return run_pipeline(ctx, params, WALK_MODE_PROBE, download_result, out, nullptr, false);
}

extern "C" int phi_gpu_index_run(phi_gpu_index_ctx *ctx, const phi_graph_view *graph, const phi_reads_view *reads, const phi_index_params *params,
                                 phi_index_result **out)
{
    if (!ctx || !out) return PHI_ERR_ARG;
    *out = nullptr;
    int rc = validate_params(ctx, params);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev[EV_START], ctx->st));
    rc = upload_async(ctx, graph, reads);                                   // copy stream: graph, then the reads in pieces; nothing waits here
    if (rc) { cudaStreamSynchronize(ctx->st_copy); ctx->have_inputs = false; ctx->n_pieces = 0; return rc; }
    CU(cudaEventRecord(ctx->ev[EV_H2D], ctx->st_copy));                     // the last read piece is the last thing to arrive
    rc = run_pipeline(ctx, params, WALK_MODE_PROBE, 1, out, nullptr, true);
    ctx->n_pieces = 0;
    if (rc) { cudaStreamSynchronize(ctx->st_copy); cudaStreamSynchronize(ctx->st); cudaStreamSynchronize(ctx->st2); }   // the caller's buffers may go away after we return
    return rc;
}

extern "C" int phi_gpu_index_sketch_walks(phi_gpu_index_ctx *ctx, const phi_graph_view *graph, const phi_index_params *params,
                                          phi_index_result **out, uint64_t **hashes_out)
{
    if (!ctx || !out || !hashes_out) return PHI_ERR_ARG;
    *out = nullptr; *hashes_out = nullptr;
    phi_reads_view none = {0, nullptr, nullptr};
    int rc = phi_gpu_index_upload(ctx, graph, &none);
    if (rc) return rc;
    return run_pipeline(ctx, params, WALK_MODE_ALL, 1, out, hashes_out, false);
}

extern "C" int phi_gpu_index_set_walk_sharing(phi_gpu_index_ctx *ctx, int chunk_shift, int share)
{
    if (!ctx) return PHI_ERR_ARG;
    if (chunk_shift < 4 || chunk_shift > 24) return ctx->fail(PHI_ERR_ARG, "chunk_shift must be in [4, 24]");
    ctx->chunk_shift = chunk_shift; ctx->dedupe = share ? 1 : 0;
    return PHI_OK;
}

extern "C" int phi_gpu_index_set_walk_region(phi_gpu_index_ctx *ctx, uint64_t coord_lo, uint64_t coord_hi)
{
    if (!ctx) return PHI_ERR_ARG;
    if (coord_lo > coord_hi) return ctx->fail(PHI_ERR_ARG, "walk region: coord_lo > coord_hi");
    ctx->own_lo = coord_lo; ctx->own_hi = coord_hi;
    return PHI_OK;
}

extern "C" int phi_gpu_index_last_sharing(const phi_gpu_index_ctx *ctx, phi_walk_sharing_stats *out)
{
    if (!ctx || !out) return PHI_ERR_ARG;
    out->chunks = ctx->n_chunks; out->active_chunks = ctx->active_chunks; out->tiles = ctx->n_tiles;
    out->unique_windows = ctx->unique_windows; out->unique_hits = ctx->unique_hits;
    return PHI_OK;
}

extern "C" int phi_gpu_index_last_times(const phi_gpu_index_ctx *ctx, phi_stage_times *out)
{
    if (!ctx || !out) return PHI_ERR_ARG;
    *out = ctx->times;
    return PHI_OK;
}

extern "C" int phi_gpu_hash128_to_64(phi_gpu_index_ctx *ctx, const uint8_t *keys, uint64_t n, int32_t len, uint64_t *out)
{
    if (!ctx || !keys || !out || len < 1 || len > 255) return ctx ? ctx->fail(PHI_ERR_ARG, "bad arguments (len must be 1..255)") : PHI_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    DevBuf dk, dout;
    CU(dk.reserve(n * len)); CU(dout.reserve(n * 8));
    CU(cudaMemcpyAsync(dk.p, keys, n * len, cudaMemcpyHostToDevice, ctx->st));
    CU(launch_hash_bytes(dk.as<uint8_t>(), n, len, dout.as<uint64_t>(), ctx->st));
    CU(cudaMemcpyAsync(out, dout.p, n * 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    dk.release(); dout.release();
    return PHI_OK;
}

