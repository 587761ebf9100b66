// Threshold filter + final ordering of the walk hits
// (/root/reference/src/ILP_index.cpp:670-722), on the device.
//
// Reference semantics (SURVEY.md §9 rule 10): per rank r, hits are grouped by their vertex list; if
// ANY group has count >= threshold * num_walks (float compare) the whole rank is dropped; otherwise
// the hits are re-emitted in std::map<std::string> order of the key "v0_v1_..._" and, inside a
// group, in insertion order (walk asc, path position asc).  The final index j inside
// Anchor_hits[r][h] is therefore (key order, position) among the hits of (r, h).
//
// Device formulation:
//   1. group table  : open addressing over (rank, vertex list) -> count, exact (lists are compared,
//                     the hash only picks the start slot).
//   2. mark         : the addition that takes a slot's count to the threshold drops its rank (inside the group count kernel).
//   3. survivors    : the hits of the representative chunks are instantiated for every member chunk (chunks.cu) in
//                     (walk, position) order, then a stable radix sort on the rank alone gives (rank, walk, position).
//   4. multi-hit fix: only (rank, walk) groups with >= 2 hits need the decimal-string key order;
//                     small groups by insertion sort in one thread, big ones by a block rank sort.
//   5. CSR          : scan of list lengths + gather.
#include "kernels.h"
#include "device_common.cuh"

namespace phi {

#define PHI_LAUNCH_CHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return e_; if (launches) ++*launches; } while (0)

constexpr uint32_t G_EMPTY = 0xFFFFFFFFu;
constexpr int SMALL_GROUP = 48;
constexpr uint32_t CSR_HIST_MAX = 8192;   // walks whose counters fit a per-block shared histogram

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// ---- probe records: what the group table compares.  (rank, number of vertices, the first PROBE_VTX vertices) of a hit in one
// 32-byte sector; longer lists (rare: k-mers over many tiny segments) continue in the vertex pool.
constexpr uint32_t PROBE_VTX = 6;
__global__ void __launch_bounds__(256) probe_build_kernel(FilterArgs A, uint4 *probe)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= A.n_hits) return;
    const int32_t *p = A.vtx_pool + A.hit_voff[i];
    const uint32_t n = A.hit_nv[i];
    uint32_t v[PROBE_VTX];
    #pragma unroll
    for (uint32_t q = 0; q < PROBE_VTX; ++q) v[q] = q < n ? (uint32_t)p[q] : 0u;
    probe[2 * i] = make_uint4(A.hit_rank[i], n, v[0], v[1]);
    probe[2 * i + 1] = make_uint4(v[2], v[3], v[4], v[5]);
}

__global__ void __launch_bounds__(256) group_count_kernel(FilterArgs A, FilterWork W)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= A.n_hits) return;
    const uint4 m0 = W.probe[2 * i], m1 = W.probe[2 * i + 1];
    const uint32_t n = m0.y;
    const int32_t *tail = n > PROBE_VTX ? A.vtx_pool + A.hit_voff[i] : nullptr;
    uint64_t hsh = mix64(m0.x + 0x9E3779B97F4A7C15ull);
    {
        const uint32_t v[PROBE_VTX] = {m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
        #pragma unroll
        for (uint32_t q = 0; q < PROBE_VTX; ++q) if (q < n) hsh = mix64(hsh ^ (uint64_t)v[q]);
        for (uint32_t q = PROBE_VTX; q < n; ++q) hsh = mix64(hsh ^ (uint64_t)(uint32_t)tail[q]);
    }
    const uint64_t mask = W.g_cap - 1;
    uint64_t slot = hsh & mask;
    for (uint64_t tries = 0; tries <= mask; ++tries) {
        uint32_t rep = W.g_slot[slot].x;
        if (rep == G_EMPTY) {
            if (W.ctr[CTR_GROUPS] * 10 > W.g_cap * 8) break;             // table too full: the host retries with a larger one
            uint32_t old = atomicCAS(&W.g_slot[slot].x, G_EMPTY, (uint32_t)i);
            if (old == G_EMPTY) { rep = (uint32_t)i; atomicAdd(&W.ctr[CTR_GROUPS], 1ull); } else rep = old;
        }
        bool same = rep == (uint32_t)i;
        if (!same) {                                                      // exact: (rank, list) against the representing record's
            const uint4 r0 = W.probe[2 * (uint64_t)rep], r1 = W.probe[2 * (uint64_t)rep + 1];
            same = r0.x == m0.x && r0.y == m0.y && r0.z == m0.z && r0.w == m0.w && r1.x == m1.x && r1.y == m1.y && r1.z == m1.z && r1.w == m1.w;
            if (same && n > PROBE_VTX) {
                const int32_t *rt = A.vtx_pool + A.hit_voff[rep];
                for (uint32_t q = PROBE_VTX; q < n; ++q) same &= rt[q] == tail[q];
            }
        }
        if (same) {
            const uint32_t wt = W.weight ? W.weight[i] : W.chunk_weight ? W.chunk_weight[A.hit_walk[i]] : 1u;
            const uint32_t before = atomicAdd(&W.g_slot[slot].y, wt);
            W.hit_slot[i] = (uint32_t)slot;
            if (W.hit_sub) W.hit_sub[i] = before;
            // anchor.second.first >= threshold * num_walks — int32 promoted to float (:698).  Counts only grow, so the group ends
            // at or above the threshold iff some addition lands there: that thread drops the rank (and counts it once).
            if (W.mark_inline && (float)(int32_t)(before + wt) >= A.thr) {
                const uint32_t r = m0.x, bit = 1u << (8 * (r & 3u));
                const uint32_t old = atomicOr((uint32_t *)(W.rank_drop + (r & ~3u)), bit);
                if (!(old & bit)) atomicAdd(&W.ctr[CTR_FILTERED], 1ull);
            }
            return;
        }
        slot = (slot + 1) & mask;
    }
    W.ctr[CTR_GROUP_OVERFLOW] = 1;
}

// ---- the -d1 statistic (/root/reference/src/ILP_index.cpp:565-606): (hash, walk) pairs sorted by hash, walks ascending inside
// a hash.  The thread at the start of a hash run counts the distinct walks of the run and bumps hist[count] (block-local
// histogram first: few distinct counts -> heavy contention otherwise).
__global__ void shared_kmer_hist_kernel(const uint64_t *hash, const uint32_t *walk, uint64_t n, uint32_t n_walks, unsigned long long *hist,
                                        unsigned long long *distinct)
{
    extern __shared__ uint32_t s_h[];
    const bool local = n_walks + 1 <= CSR_HIST_MAX;
    if (local) { for (uint32_t i = threadIdx.x; i <= n_walks; i += blockDim.x) s_h[i] = 0; __syncthreads(); }
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t cnt = 0;
    if (i < n && (i == 0 || hash[i] != hash[i - 1])) {
        const uint64_t h = hash[i];
        cnt = 1;
        for (uint64_t j = i + 1; j < n && hash[j] == h; ++j) cnt += walk[j] != walk[j - 1];
        if (local) atomicAdd(&s_h[cnt], 1u); else atomicAdd(&hist[cnt], 1ull);
    }
    const uint32_t heads = __syncthreads_count(cnt != 0);
    if (threadIdx.x == 0 && heads) atomicAdd(distinct, (unsigned long long)heads);
    if (local) for (uint32_t q = threadIdx.x; q <= n_walks; q += blockDim.x) if (s_h[q]) atomicAdd(&hist[q], (unsigned long long)s_h[q]);
}

cudaError_t filter_shared_kmer_hist(const uint64_t *hash, const uint32_t *walk, uint64_t n, uint32_t n_walks, unsigned long long *hist,
                                    unsigned long long *distinct, cudaStream_t st, uint64_t *launches)
{
    if (!n) return cudaSuccess;
    const size_t smem = n_walks + 1 <= CSR_HIST_MAX ? ((size_t)n_walks + 1) * 4 : 0;
    shared_kmer_hist_kernel<<<(unsigned)((n + 255) / 256), 256, smem, st>>>(hash, walk, n, n_walks, hist, distinct);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t filter_build_probe(const FilterArgs &A, uint4 *probe, cudaStream_t st, uint64_t *launches)
{
    if (!A.n_hits) return cudaSuccess;
    probe_build_kernel<<<(unsigned)((A.n_hits + 255) / 256), 256, 0, st>>>(A, probe);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t filter_count_groups(const FilterArgs &A, const FilterWork &W, cudaStream_t st, uint64_t *launches)
{
    if (!A.n_hits) return cudaSuccess;
    group_count_kernel<<<(unsigned)((A.n_hits + 255) / 256), 256, 0, st>>>(A, W);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}
__global__ void emit_rank_keys_kernel(FilterArgs A, uint32_t *keys, uint32_t *vals)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < A.n_hits) { keys[i] = A.hit_rank[i]; vals[i] = (uint32_t)i; }
}

// The records arrive in (walk, position) order: a stable sort of (u32 rank, u32 record) pairs on the rank gives the final
// (rank, walk, position) order.  Wide digits; the key arrays live in keys_a / keys_b.
cudaError_t filter_sort_records(const FilterArgs &A, const FilterWork &W, cudaStream_t st, uint64_t *launches)
{
    const uint64_t n = A.n_hits;
    if (!n) return cudaSuccess;
    emit_rank_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(A, (uint32_t *)W.keys_a, W.vals_a);
    PHI_LAUNCH_CHECK();
    return radix_sort_u32((uint32_t *)W.keys_a, (uint32_t *)W.keys_b, W.vals_a, W.vals_b, n, A.rank_bits, W.sort_scratch, st, launches);
}

// ---- decimal-string order of "v0_v1_..._" keys ('_' sorts after every digit)
__device__ __forceinline__ int ndigits(uint32_t v)
{
    int d = 1;
    while (v >= 10) { v /= 10; ++d; }
    return d;
}
__device__ __forceinline__ int cmp_dec(uint32_t a, uint32_t b)        // order of to_string(a)+"_" vs to_string(b)+"_"
{
    if (a == b) return 0;
    int da = ndigits(a), db = ndigits(b);
    if (da == db) return a < b ? -1 : 1;
    if (da < db) {                                                    // a shorter: compare a with the first da digits of b
        uint32_t bp = b; for (int i = 0; i < db - da; ++i) bp /= 10;
        if (a != bp) return a < bp ? -1 : 1;
        return 1;                                                     // a is a proper prefix: a's '_' meets a digit of b -> a sorts after
    }
    uint32_t ap = a; for (int i = 0; i < da - db; ++i) ap /= 10;
    if (ap != b) return ap < b ? -1 : 1;
    return -1;
}
__device__ int cmp_list(const FilterArgs &A, uint32_t x, uint32_t y)
{
    uint32_t nx = A.hit_nv[x], ny = A.hit_nv[y];
    const int32_t *px = A.vtx_pool + A.hit_voff[x], *py = A.vtx_pool + A.hit_voff[y];
    uint32_t n = nx < ny ? nx : ny;
    for (uint32_t i = 0; i < n; ++i) { int c = cmp_dec((uint32_t)px[i], (uint32_t)py[i]); if (c) return c; }
    return nx == ny ? 0 : nx < ny ? -1 : 1;                           // proper prefix string sorts first
}

__device__ __forceinline__ bool same_rw(const FilterArgs &A, uint32_t x, uint32_t y, int by_walk)
{
    return A.hit_rank[x] == A.hit_rank[y] && (!by_walk || A.hit_walk[x] == A.hit_walk[y]);
}

// order[] = hit ids sorted by (rank, walk, position).  Each thread owning the head of a (rank, walk) group with >= 2 members
// re-orders it by (key string, position): stable insertion sort for small groups, deferral for big ones.
__global__ void fix_multi_kernel(FilterArgs A, uint32_t *order, uint64_t n, int by_walk, uint32_t *big_list, uint32_t big_cap, unsigned long long *ctr)
{
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t me = order[j];
    if (j > 0 && same_rw(A, order[j - 1], me, by_walk)) return;               // not a head
    if (j + 1 >= n || !same_rw(A, order[j + 1], me, by_walk)) return;         // singleton
    uint64_t e = j + 2;
    while (e < n && same_rw(A, order[e], me, by_walk)) ++e;
    uint64_t len = e - j;
    if (len > SMALL_GROUP) {
        unsigned long long slot = atomicAdd(&ctr[CTR_BIG_GROUPS], 1ull);
        if (slot < big_cap) { big_list[2 * slot] = (uint32_t)j; big_list[2 * slot + 1] = (uint32_t)len; }
        return;
    }
    for (uint64_t a = j + 1; a < e; ++a) {
        uint32_t x = order[a]; uint64_t b = a;
        while (b > j && cmp_list(A, order[b - 1], x) > 0) { order[b] = order[b - 1]; --b; }
        order[b] = x;
    }
}

// one block per big group: rank sort (stable: ties keep the incoming position order)
__global__ void fix_big_kernel(FilterArgs A, uint32_t *order, uint32_t *tmp, const uint32_t *big_list, uint32_t big_cap, const unsigned long long *ctr)
{
    const uint32_t n_big = (uint32_t)min((unsigned long long)big_cap, ctr[CTR_BIG_GROUPS]);
    for (uint32_t g = blockIdx.x; g < n_big; g += gridDim.x) {
        const uint32_t j0 = big_list[2 * g], len = big_list[2 * g + 1];
        for (uint32_t a = threadIdx.x; a < len; a += blockDim.x) {
            uint32_t x = order[j0 + a], pos = 0;
            for (uint32_t b = 0; b < len; ++b) {
                if (b == a) continue;
                int c = cmp_list(A, order[j0 + b], x);
                pos += (c < 0) || (c == 0 && b < a);
            }
            tmp[j0 + pos] = x;
        }
        __syncthreads();
        for (uint32_t a = threadIdx.x; a < len; a += blockDim.x) order[j0 + a] = tmp[j0 + a];
        __syncthreads();
    }
}

cudaError_t filter_fix_multi(const FilterArgs &A, uint32_t *order, uint64_t n_surv, int by_walk, uint32_t *big_list, uint32_t big_cap,
                             unsigned long long *ctr, cudaStream_t st, uint64_t *launches)
{
    if (n_surv < 2) return cudaSuccess;
    fix_multi_kernel<<<(unsigned)((n_surv + 127) / 128), 128, 0, st>>>(A, order, n_surv, by_walk, big_list, big_cap, ctr);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}
cudaError_t filter_fix_big(const FilterArgs &A, uint32_t *order, uint32_t *tmp, const uint32_t *big_list, uint32_t big_cap,
                           const unsigned long long *ctr, cudaStream_t st, uint64_t *launches)
{
    fix_big_kernel<<<148, 256, 0, st>>>(A, order, tmp, big_list, big_cap, ctr);     // usually nothing to do: the blocks leave at once
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

__global__ void csr_sizes_kernel(FilterArgs A, const uint32_t *order, uint64_t n, uint32_t *nv_out, uint8_t *anchor_len)
{
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j < n) { uint8_t nv = A.hit_nv[order[j]]; nv_out[j] = nv; anchor_len[j] = nv; }
}

// final arrays of the result: walk of every anchor, vertex lists back to back, first anchor of every rank (rank_off, optional)
__global__ void csr_fill_kernel(FilterArgs A, const uint32_t *order, uint64_t n, const uint64_t *anchor_off, uint64_t *rank_off,
                                int32_t *anchor_walk, int32_t *anchor_vtx, unsigned long long *anchors_per_walk, uint32_t n_walks_out)
{
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t wl = 0xFFFFFFFFu;
    if (j < n) {
        uint32_t x = order[j];
        if (rank_off) {                                               // anchors are sorted by rank: rank r starts where it first appears
            const int64_t r = A.hit_rank[x], rp = j ? (int64_t)A.hit_rank[order[j - 1]] : -1;
            for (int64_t q = rp + 1; q <= r; ++q) rank_off[q] = j;
            if (j == n - 1) for (int64_t q = r + 1; q <= (int64_t)A.n_ranks; ++q) rank_off[q] = n;
        }
        anchor_walk[j] = (int32_t)A.hit_walk[x];
        const int32_t *p = A.vtx_pool + A.hit_voff[x];
        uint64_t o = anchor_off[j]; uint32_t nv = A.hit_nv[x];
        for (uint32_t i = 0; i < nv; ++i) anchor_vtx[o + i] = p[i];
        wl = A.hit_walk[x];
    }
    // per-walk counts: block-local shared histogram (few distinct walks -> heavy global contention otherwise)
    extern __shared__ uint32_t s_hist[];
    if (n_walks_out <= CSR_HIST_MAX) {
        for (uint32_t i = threadIdx.x; i < n_walks_out; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        if (wl != 0xFFFFFFFFu) atomicAdd(&s_hist[wl], 1u);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_walks_out; i += blockDim.x)
            if (s_hist[i]) atomicAdd(&anchors_per_walk[i], (unsigned long long)s_hist[i]);
    } else {
        uint32_t peers = __match_any_sync(0xFFFFFFFFu, wl);
        if (wl != 0xFFFFFFFFu && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&anchors_per_walk[wl], (unsigned long long)__popc(peers));
    }
}

cudaError_t filter_csr_sizes(const FilterArgs &A, const uint32_t *order, uint64_t n_surv, uint32_t *nv_out, uint8_t *anchor_len, cudaStream_t st, uint64_t *launches)
{
    if (!n_surv) return cudaSuccess;
    csr_sizes_kernel<<<(unsigned)((n_surv + 255) / 256), 256, 0, st>>>(A, order, n_surv, nv_out, anchor_len);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}
cudaError_t filter_csr_fill(const FilterArgs &A, const uint32_t *order, uint64_t n_surv, const uint64_t *anchor_off, uint64_t *rank_off,
                            int32_t *anchor_walk, int32_t *anchor_vtx, unsigned long long *anchors_per_walk,
                            uint32_t n_walks_out, cudaStream_t st, uint64_t *launches)
{
    if (!n_surv) return cudaSuccess;
    const size_t hist_bytes = n_walks_out <= CSR_HIST_MAX ? (size_t)n_walks_out * 4 : 0;
    csr_fill_kernel<<<(unsigned)((n_surv + 1023) / 1024), 1024, hist_bytes, st>>>(A, order, n_surv, anchor_off, rank_off, anchor_walk, anchor_vtx,
                                                                     anchors_per_walk, n_walks_out);
    PHI_LAUNCH_CHECK();
    return cudaSuccess;
}

}  // namespace phi
