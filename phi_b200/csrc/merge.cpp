// phi_index_result_merge: the per-GPU parts of a multi-GPU run -> one result in the reference's order (host only).
//
// Every part holds, for all hash ranks, the surviving groups (rank, vertex list) found on ITS GPU, in the reference's key order
// (/root/reference/src/ILP_index.cpp:680-709: std::map<std::string> over "v0_v1_..._"), with ascending member walks.  The groups
// of one rank are merged in key order; a group that occurs in several parts (its occurrences were owned by several GPUs) gets
// the union of the member lists.  Per-walk counters and n_filtered are sums over the parts; the spectrum comes from the part
// that carries it (rank 0).  With one part this is a copy.
// Big merges run on several host threads over blocks of hash ranks (count pass, prefix sum, write pass): PHI_MERGE_THREADS.
#include "result_box.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <sys/mman.h>

namespace {

std::string key_of(const int32_t *v, uint32_t n)
{
    std::string s;
    char buf[16];
    for (uint32_t i = 0; i < n; ++i) { int len = snprintf(buf, sizeof buf, "%d_", v[i]); s.append(buf, (size_t)len); }
    return s;
}

template <class T> T *heap_array(ResultBox *b, uint64_t n)
{
    // big arrays: 2 MB aligned and marked for transparent huge pages — the first touch of the output is a large share of the merge
    // with 4 KB pages (one fault per page, and the faults of all threads meet in the kernel's address-space lock)
    const size_t bytes = (size_t)(n ? n : 1) * sizeof(T);
    T *p = nullptr;
    if (bytes >= (4u << 20)) {
        void *q = nullptr;
        if (posix_memalign(&q, 2u << 20, bytes) == 0) { p = (T *)q; madvise(q, bytes, MADV_HUGEPAGE); }
    }
    if (!p) p = (T *)malloc(bytes);
    if (p) { b->bufs[b->nbufs].p = p; b->bufs[b->nbufs].cap = n * sizeof(T); b->nbufs++; }
    return p;
}

struct Totals { uint64_t ng, nm, nv; };              // groups, members, vertices written so far

// PHI_MERGE_THREADS, else up to 16 of the machine's threads; small merges are not worth a thread start.
int merge_threads(uint64_t work)
{
    if (const char *e = getenv("PHI_MERGE_THREADS")) { const int t = atoi(e); if (t > 0) return std::min(t, 64); }
    if (work < (1u << 20)) return 1;
    return (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
}

// fn(block) for every block in [0, n_blocks), blocks handed out to n_threads threads through a shared counter.
template <class F> void parallel_blocks(int n_blocks, int n_threads, F fn)
{
    n_threads = std::min(n_threads, n_blocks);
    if (n_threads <= 1) { for (int b = 0; b < n_blocks; ++b) fn(b); return; }
    std::atomic<int> next(0);
    auto work = [&]() { for (int b; (b = next.fetch_add(1)) < n_blocks;) fn(b); };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (std::thread &t : pool) t.join();
}

}  // namespace

extern "C" int phi_index_result_merge(const phi_index_result *const *parts, int n_parts, phi_index_result **out)
{
    if (!parts || n_parts < 1 || !out) return PHI_ERR_ARG;
    *out = nullptr;
    const phi_index_result *p0 = parts[0];
    const int32_t NS = p0->count_sp_r; const uint32_t NW = p0->n_walks;
    uint64_t tot_groups = 0, tot_members = 0, tot_vtx = 0;
    const uint64_t *spectrum = nullptr;
    for (int p = 0; p < n_parts; ++p) {
        const phi_index_result *r = parts[p];
        if (!r || r->count_sp_r != NS || r->n_walks != NW) return PHI_ERR_ARG;
        if (r->n_groups && (!r->rank_off || !r->group_len || !r->group_member_off || (!r->member_walk16 && !r->member_walk32))) return PHI_ERR_ARG;
        if (NS && !r->rank_off) return PHI_ERR_ARG;
        tot_groups += r->n_groups; tot_members += r->n_anchors; tot_vtx += r->n_group_vtx;
        if (!spectrum && r->spectrum) spectrum = r->spectrum;
    }
    if (tot_groups >= (1ull << 32) || tot_members >= (1ull << 32)) return PHI_ERR_UNSUPPORTED;   // the ABI's offsets are u32
    ResultBox *b = (ResultBox *)calloc(1, sizeof(ResultBox));
    if (!b) return PHI_ERR_NOMEM;
    b->heap = 1;
    phi_index_result *m = &b->pub;
    const bool w16 = NW <= 65536;
    uint64_t *o_spec = heap_array<uint64_t>(b, spectrum ? (uint64_t)NS : 0);
    uint32_t *o_rank_off = heap_array<uint32_t>(b, (uint64_t)NS + 1);
    uint8_t *o_len = heap_array<uint8_t>(b, tot_groups);
    int32_t *o_vtx = heap_array<int32_t>(b, tot_vtx);
    uint32_t *o_moff = heap_array<uint32_t>(b, tot_groups + 1);
    uint16_t *o_w16 = w16 ? heap_array<uint16_t>(b, tot_members) : nullptr;
    int32_t *o_w32 = w16 ? nullptr : heap_array<int32_t>(b, tot_members);
    uint64_t *o_mpw = heap_array<uint64_t>(b, NW), *o_apw = heap_array<uint64_t>(b, NW);
    if (!o_spec || !o_rank_off || !o_len || !o_vtx || !o_moff || (w16 ? !o_w16 : !o_w32) || !o_mpw || !o_apw) { phi_gpu_index_result_free(m); return PHI_ERR_NOMEM; }
    memset(o_mpw, 0, (size_t)NW * 8); memset(o_apw, 0, (size_t)NW * 8);

    const bool times = getenv("PHI_MERGE_TIMES") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    // vertex offset of every group of every part (group_len is u8: the lists lie back to back)
    std::vector<std::vector<uint64_t>> gvoff(n_parts);
    parallel_blocks(n_parts, n_parts, [&](int p) {
        const phi_index_result *r = parts[p];
        gvoff[p].resize(r->n_groups + 1);
        uint64_t run = 0;
        for (uint64_t g = 0; g < r->n_groups; ++g) { gvoff[p][g] = run; run += r->group_len[g]; }
        gvoff[p][r->n_groups] = run;
    });
    // The hash ranks are cut into blocks; a first pass counts what every block will write (groups, members, vertices: a group held
    // by several parts is written once, so the totals are not the sums of the parts), a prefix sum places the blocks, a second pass
    // writes them.  Both passes run the blocks on several threads: the merge is memory traffic (rank_off of every part, ~2 bytes
    // per anchor of member walks) plus the first touch of the output pages.
    const int n_threads = merge_threads((uint64_t)NS * (uint64_t)n_parts + tot_members);
    const int n_blocks = n_threads == 1 ? 1 : n_threads * 4;
    std::vector<Totals> at((size_t)n_blocks + 1);
    auto rank_lo = [&](int blk) { return (int32_t)((int64_t)NS * blk / n_blocks); };
    auto pass = [&](int blk, bool write) {
        Totals t = write ? at[blk] : Totals{0, 0, 0};
        uint64_t &ng = t.ng, &nm = t.nm, &nv = t.nv;
        auto member = [&](const phi_index_result *r, uint64_t i) -> uint32_t { return r->member_walk16 ? (uint32_t)r->member_walk16[i] : (uint32_t)r->member_walk32[i]; };
        auto put_member = [&](uint32_t wk) { if (w16) o_w16[nm] = (uint16_t)wk; else o_w32[nm] = (int32_t)wk; };
        std::vector<int> live; std::vector<uint32_t> cur(n_parts), end(n_parts);
        std::vector<const uint32_t *> ro(n_parts);
        for (int p = 0; p < n_parts; ++p) ro[p] = parts[p]->rank_off;
        std::vector<std::string> keys(n_parts);
        std::vector<uint32_t> merged;
        const int32_t rk0 = rank_lo(blk), rk1 = rank_lo(blk + 1);
        for (int p = 0; p < n_parts; ++p) end[p] = rk0 < rk1 ? ro[p][rk0] : 0;
        for (int32_t rk = rk0; rk < rk1; ++rk) {
            if (write) o_rank_off[rk] = (uint32_t)ng;
            // most ranks have no surviving group in any part, nearly all others in exactly one: find those with one load per part
            int n_live = 0, one = -1;
            for (int p = 0; p < n_parts; ++p) {
                cur[p] = end[p]; end[p] = ro[p][rk + 1];
                if (cur[p] < end[p]) { ++n_live; one = p; }
            }
            if (!n_live) continue;
            if (n_live == 1) {                                            // the usual case: all groups of this rank come from one GPU
                // block copies: the groups, their vertex lists and their member walks lie back to back in the part
                const int p = one; const phi_index_result *r = parts[p];
                const uint32_t g0 = cur[p], g1 = end[p];
                const uint32_t m0 = r->group_member_off[g0], m1 = r->group_member_off[g1];
                const uint64_t v0 = gvoff[p][g0], v1 = gvoff[p][g1];
                if (write) {
                    memcpy(o_len + ng, r->group_len + g0, (size_t)(g1 - g0));
                    memcpy(o_vtx + nv, r->group_vtx + v0, (size_t)(v1 - v0) * 4);
                    for (uint32_t g = g0; g < g1; ++g) o_moff[ng + (g - g0)] = (uint32_t)(nm + (r->group_member_off[g] - m0));
                    if (w16 && r->member_walk16) memcpy(o_w16 + nm, r->member_walk16 + m0, (size_t)(m1 - m0) * 2);
                    else if (!w16 && r->member_walk32) memcpy(o_w32 + nm, r->member_walk32 + m0, (size_t)(m1 - m0) * 4);
                    else for (uint32_t i = m0; i < m1; ++i) { if (w16) o_w16[nm + (i - m0)] = (uint16_t)member(r, i); else o_w32[nm + (i - m0)] = (int32_t)member(r, i); }
                }
                ng += g1 - g0; nv += v1 - v0; nm += m1 - m0;
                continue;
            }
            live.clear();
            for (int p = 0; p < n_parts; ++p) if (cur[p] < end[p]) live.push_back(p);
            for (int p : live) keys[p] = key_of(parts[p]->group_vtx + gvoff[p][cur[p]], parts[p]->group_len[cur[p]]);
            while (!live.empty()) {
                int best = live[0];
                for (int p : live) if (keys[p] < keys[best]) best = p;
                const std::string key = keys[best];
                const phi_index_result *rb = parts[best];
                const uint8_t len = rb->group_len[cur[best]];
                if (write) {
                    o_len[ng] = len; o_moff[ng] = (uint32_t)nm;
                    memcpy(o_vtx + nv, rb->group_vtx + gvoff[best][cur[best]], (size_t)len * 4);
                }
                nv += len;
                // members: union over the parts that hold this key, ascending (every part's list ascends: merge)
                merged.clear();
                for (size_t li = 0; li < live.size();) {
                    const int p = live[li];
                    if (keys[p] != key) { ++li; continue; }
                    const phi_index_result *r = parts[p];
                    const uint32_t g = cur[p];
                    const size_t old = merged.size();
                    for (uint32_t i = r->group_member_off[g]; i < r->group_member_off[g + 1]; ++i) merged.push_back(member(r, i));
                    std::inplace_merge(merged.begin(), merged.begin() + old, merged.end());
                    if (++cur[p] < end[p]) { keys[p] = key_of(r->group_vtx + gvoff[p][cur[p]], r->group_len[cur[p]]); ++li; }
                    else live.erase(live.begin() + li);
                }
                for (uint32_t wk : merged) { if (write) put_member(wk); ++nm; }
                ++ng;
            }
        }
        if (!write) at[blk + 1] = t;
    };
    const double t_pre = now();
    parallel_blocks(n_blocks, n_threads, [&](int blk) { pass(blk, false); });
    const double t_count = now();
    at[0] = Totals{0, 0, 0};
    for (int blk = 1; blk <= n_blocks; ++blk) { at[blk].ng += at[blk - 1].ng; at[blk].nm += at[blk - 1].nm; at[blk].nv += at[blk - 1].nv; }
    const uint64_t ng = at[n_blocks].ng, nm = at[n_blocks].nm, nv = at[n_blocks].nv;
    parallel_blocks(n_blocks, n_threads, [&](int blk) {
        pass(blk, true);
        if (spectrum && NS) {                                             // the spectrum travels block by block with the ranks
            const int32_t a = rank_lo(blk), z = rank_lo(blk + 1);
            memcpy(o_spec + a, spectrum + a, (size_t)(z - a) * 8);
        }
    });
    o_rank_off[NS] = (uint32_t)ng; o_moff[ng] = (uint32_t)nm;
    if (times) fprintf(stderr, "[phi_index_result_merge] %d parts, %d threads: offsets %.1f ms, count pass %.1f ms, write pass %.1f ms\n", n_parts, n_threads, t_pre - t_start, t_count - t_pre, now() - t_count);
    for (int p = 0; p < n_parts; ++p) {
        const phi_index_result *r = parts[p];
        for (uint32_t h = 0; h < NW; ++h) { if (r->minimizers_per_walk) o_mpw[h] += r->minimizers_per_walk[h]; if (r->anchors_per_walk) o_apw[h] += r->anchors_per_walk[h]; }
        m->n_filtered += r->n_filtered;
        m->read_kmer_positions += r->read_kmer_positions; m->path_kmer_positions += r->path_kmer_positions;
        m->read_minimizers_emitted += r->read_minimizers_emitted; m->path_minimizers_emitted += r->path_minimizers_emitted;
        m->path_hits += r->path_hits;
    }
    for (int p = 0; p < n_parts; ++p) {                                  // the -d1 statistic is global already (every part carries the same)
        const phi_index_result *r = parts[p];
        if (!r->shared_kmer_hist) continue;
        uint64_t *o_hist = heap_array<uint64_t>(b, (uint64_t)NW + 1);
        if (!o_hist) { phi_gpu_index_result_free(m); return PHI_ERR_NOMEM; }
        memcpy(o_hist, r->shared_kmer_hist, ((size_t)NW + 1) * 8);
        m->shared_kmer_hist = o_hist; m->n_walk_kmers = r->n_walk_kmers;
        break;
    }
    m->count_sp_r = NS; m->n_walks = NW; m->n_anchors = nm; m->n_groups = ng; m->n_group_vtx = nv;
    m->spectrum = spectrum ? o_spec : nullptr; m->rank_off = o_rank_off; m->group_len = o_len; m->group_vtx = o_vtx; m->group_member_off = o_moff;
    m->member_walk16 = o_w16; m->member_walk32 = o_w32; m->minimizers_per_walk = o_mpw; m->anchors_per_walk = o_apw;
    *out = m;
    return PHI_OK;
}
