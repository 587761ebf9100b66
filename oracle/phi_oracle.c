/*
 * TEST INFRASTRUCTURE ONLY — see phi_oracle.h.  Plain C, byte strings, no
 * 2-bit tricks: this file restates the reference's algorithm in the most
 * literal form that still finishes in seconds, so that it is an independent
 * check on the CUDA path (which uses packed 2-bit k-mers).
 *
 * Each function cites the reference lines it follows (/root/reference/src/...).
 */
#include "phi_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------- MurmurHash3
 * MurmurHash3.cpp:39-42 (rotl64), :81-90 (fmix64), :255-332 (x64_128).          */
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}
static inline uint64_t load_le64(const uint8_t *p)
{
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}

void phi_oracle_murmur3_x64_128(const uint8_t *data, int32_t len, uint32_t seed, uint64_t out[2])
{
    const int nblocks = len / 16;
    uint64_t h1 = seed, h2 = seed;
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    for (int i = 0; i < nblocks; ++i) {                         /* body :270-282 */
        uint64_t k1 = load_le64(data + 16 * i), k2 = load_le64(data + 16 * i + 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + nblocks * 16;                  /* tail :287-314 */
    uint64_t k1 = 0, k2 = 0;
    int rem = len & 15;
    for (int i = rem - 1; i >= 8; --i) k2 ^= (uint64_t)tail[i] << (8 * (i - 8));
    if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (int i = (rem > 8 ? 8 : rem) - 1; i >= 0; --i) k1 ^= (uint64_t)tail[i] << (8 * i);
    if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;                   /* finalization :319-331 */
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

/* ILP_index.cpp:10-18 */
uint64_t phi_oracle_hash128_to_64(const uint8_t *key, int32_t len)
{
    uint64_t o[2];
    phi_oracle_murmur3_x64_128(key, len, 0, o);
    return o[0] ^ o[1];
}

/* ------------------------------------------------------------ string helpers */
/* ::toupper in the C locale (ILP_index.cpp:369, :449) */
static inline uint8_t up(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

/* reverse_strand, ILP_index.cpp:330-357: reverse; A<->T, C<->G (either case in, upper case out); other bytes verbatim */
static void reverse_strand(const uint8_t *s, int k, uint8_t *out)
{
    for (int i = 0; i < k; ++i) {
        uint8_t c = s[k - 1 - i], r;
        if (c == 'A' || c == 'a') r = 'T';
        else if (c == 'T' || c == 't') r = 'A';
        else if (c == 'C' || c == 'c') r = 'G';
        else if (c == 'G' || c == 'g') r = 'C';
        else r = c;
        out[i] = r;
    }
}

/*
 * The minimizer scan shared by index_kmers (ILP_index.cpp:383-442) and
 * compute_hashes (:455-490).  `seq` is already upper-cased, length n.
 * Calls emit(ctx, hash, best_start_idx, i) for every emitted minimizer in order; i is the start of the LAST k-mer of the
 * window that emitted it (the loop variable of :388 at :413).
 */
typedef void (*emit_fn)(void *ctx, uint64_t hash, int64_t pos, int64_t win_end);

static void minimizer_scan(const uint8_t *seq, int64_t n, int k, int w, emit_fn emit, void *ctx)
{
    if (n < (int64_t)w + k - 1) return;                         /* :372, :453 */
    /* monotone deque of (canonical k-mer string, position) — at most w live entries */
    int cap = w + 1;
    uint8_t *dq_str = (uint8_t *)malloc((size_t)cap * k);
    int64_t *dq_pos = (int64_t *)malloc((size_t)cap * sizeof(int64_t));
    uint8_t *rev = (uint8_t *)malloc(k);
    int head = 0, size = 0;                                     /* ring buffer */
    uint64_t prev_hash = UINT64_MAX;                            /* :383, :455 */
    for (int64_t i = 0; i <= n - k; ++i) {                      /* :388, :460 */
        const uint8_t *fwd = seq + i;
        reverse_strand(fwd, k, rev);                            /* :391, :463 */
        const uint8_t *mn = memcmp(rev, fwd, k) < 0 ? rev : fwd; /* std::min(fwd, rev) :394 */
        while (size > 0) {                                      /* pop_back while back >= cur :397 */
            int b = (head + size - 1) % cap;
            if (memcmp(dq_str + (size_t)b * k, mn, k) >= 0) --size; else break;
        }
        int t = (head + size) % cap;                            /* emplace_back :402 */
        memcpy(dq_str + (size_t)t * k, mn, k); dq_pos[t] = i; ++size;
        if (size > 0 && dq_pos[head] <= i - w) { head = (head + 1) % cap; --size; }  /* :405-407 */
        if (i >= w - 1) {                                       /* :410 */
            uint64_t h = phi_oracle_hash128_to_64(dq_str + (size_t)head * k, k);      /* :412 */
            if (h != prev_hash) { prev_hash = h; emit(ctx, h, dq_pos[head], i); }     /* :413-414 */
        }
    }
    free(dq_str); free(dq_pos); free(rev);
}

/* ------------------------------------------------------------------ vectors */
typedef struct { uint64_t *a; size_t n, m; } vec_u64;
typedef struct { int32_t *a; size_t n, m; } vec_i32;
static void push_u64(vec_u64 *v, uint64_t x)
{
    if (v->n == v->m) { v->m = v->m ? v->m * 2 : 1024; v->a = (uint64_t *)realloc(v->a, v->m * 8); }
    v->a[v->n++] = x;
}
static void push_i32(vec_i32 *v, int32_t x)
{
    if (v->n == v->m) { v->m = v->m ? v->m * 2 : 1024; v->a = (int32_t *)realloc(v->a, v->m * 4); }
    v->a[v->n++] = x;
}
static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* ------------------------------------------------------------------- reads */
static void emit_hash_only(void *ctx, uint64_t h, int64_t pos, int64_t win_end) { (void)pos; (void)win_end; push_u64((vec_u64 *)ctx, h); }

/* compute_hashes, ILP_index.cpp:447-493 */
int64_t phi_oracle_read_hashes(const uint8_t *read, uint64_t len, int32_t k, int32_t w, uint64_t **out)
{
    uint8_t *s = (uint8_t *)malloc(len ? len : 1);
    for (uint64_t i = 0; i < len; ++i) s[i] = up(read[i]);       /* :449 */
    vec_u64 v = {0, 0, 0};
    minimizer_scan(s, (int64_t)len, k, w, emit_hash_only, &v);
    free(s);
    if (v.n) qsort(v.a, v.n, 8, cmp_u64);                        /* std::set: sorted, distinct */
    size_t m = 0;
    for (size_t i = 0; i < v.n; ++i) if (i == 0 || v.a[i] != v.a[i - 1]) v.a[m++] = v.a[i];
    *out = v.a;
    return (int64_t)m;
}

/* ------------------------------------------------------------------- walks */
typedef struct {
    vec_u64 hash;      /* per minimizer */
    vec_u64 voff;      /* per minimizer + 1 */
    vec_i32 vtx;
    vec_i32 own;       /* per minimizer: the vertex under the start of the last k-mer of the window that emitted it */
    const int32_t *idx_vtx_map;
    const int32_t *top_order_map;
    int k;
} walk_sketch;

/* anchor construction, ILP_index.cpp:416-439 */
static void emit_walk_min(void *ctx, uint64_t h, int64_t pos, int64_t win_end)
{
    walk_sketch *ws = (walk_sketch *)ctx;
    push_i32(&ws->own, ws->idx_vtx_map[win_end]);
    int32_t uniq[256]; int nu = 0;
    for (int j = 0; j < ws->k; ++j) {                            /* :424-430 distinct, first-seen order */
        int32_t v = ws->idx_vtx_map[pos + j];
        int seen = 0;
        for (int q = 0; q < nu; ++q) if (uniq[q] == v) { seen = 1; break; }
        if (!seen) uniq[nu++] = v;
    }
    /* :433-435 sort by top_order_map (insertion sort; ties cannot occur for a valid topological order) */
    for (int a = 1; a < nu; ++a) {
        int32_t x = uniq[a]; int b = a - 1;
        while (b >= 0 && ws->top_order_map[uniq[b]] > ws->top_order_map[x]) { uniq[b + 1] = uniq[b]; --b; }
        uniq[b + 1] = x;
    }
    push_u64(&ws->hash, h);
    for (int q = 0; q < nu; ++q) push_i32(&ws->vtx, uniq[q]);
    push_u64(&ws->voff, ws->vtx.n);
}

/* index_kmers(hap), ILP_index.cpp:359-445 */
static void sketch_one_walk(const phi_graph_view *g, uint32_t h, int k, int w, walk_sketch *ws)
{
    memset(ws, 0, sizeof(*ws));
    push_u64(&ws->voff, 0);
    uint64_t len = 0;
    for (uint64_t s = g->walk_off[h]; s < g->walk_off[h + 1]; ++s) {
        uint32_t v = g->walk_vtx[s];
        len += g->seg_off[v + 1] - g->seg_off[v];
    }
    uint8_t *hap = (uint8_t *)malloc(len ? len : 1);
    int32_t *map = (int32_t *)malloc((len ? len : 1) * sizeof(int32_t));
    uint64_t p = 0;
    for (uint64_t s = g->walk_off[h]; s < g->walk_off[h + 1]; ++s) {   /* :364-366, :375-381 */
        uint32_t v = g->walk_vtx[s];
        for (uint64_t b = g->seg_off[v]; b < g->seg_off[v + 1]; ++b) { hap[p] = up(g->seg_bases[b]); map[p] = (int32_t)v; ++p; }
    }
    ws->idx_vtx_map = map; ws->top_order_map = g->top_order_map; ws->k = k;
    minimizer_scan(hap, (int64_t)len, k, w, emit_walk_min, ws);
    free(hap); free(map);
    ws->idx_vtx_map = 0;
}

static uint64_t kmer_positions(uint64_t len, int k, int w)
{
    return len >= (uint64_t)(w + k - 1) ? len - k + 1 : 0;
}

static int check_params(const phi_index_params *p)
{
    return p && p->k >= 1 && p->k <= 255 && p->w >= 1;
}

/* public layout of the anchors (include/phi_gpu_index.h): per-rank offsets + per-anchor list lengths instead of the
 * per-anchor rank / offset arrays this file works with; consumes (frees) arank and aoff */
static void to_compact(phi_oracle_result *r, int32_t *arank, uint64_t *aoff, int32_t n_ranks)
{
    const uint64_t na = r->n_anchors;
    uint8_t *len = (uint8_t *)malloc(na ? na : 1);
    for (uint64_t a = 0; a < na; ++a) len[a] = (uint8_t)(aoff[a + 1] - aoff[a]);
    r->anchor_len = len;
    r->rank_off = NULL;
    if (n_ranks > 0) {
        uint64_t *ro = (uint64_t *)calloc((size_t)n_ranks + 1, 8);
        for (uint64_t a = 0; a < na; ++a) ro[arank[a] + 1]++;
        for (int32_t q = 0; q < n_ranks; ++q) ro[q + 1] += ro[q];
        r->rank_off = ro;
    }
    free(arank); free(aoff);
}

int phi_oracle_sketch_walks(const phi_graph_view *g, const phi_index_params *prm, int n_threads,
                            phi_oracle_result **out, uint64_t **hashes_out)
{
    return phi_oracle_sketch_walks_owner(g, prm, n_threads, out, hashes_out, NULL);
}

int phi_oracle_sketch_walks_owner(const phi_graph_view *g, const phi_index_params *prm, int n_threads,
                                  phi_oracle_result **out, uint64_t **hashes_out, int32_t **owner_vtx_out)
{
    if (!g || !out || !hashes_out || !check_params(prm)) return PHI_ERR_ARG;
    uint32_t H = g->n_walks;
    walk_sketch *ws = (walk_sketch *)calloc(H ? H : 1, sizeof(walk_sketch));
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#endif
    #pragma omp parallel for num_threads(n_threads) schedule(dynamic, 1)
    for (int64_t h = 0; h < (int64_t)H; ++h) sketch_one_walk(g, (uint32_t)h, prm->k, prm->w, &ws[h]);

    phi_oracle_result *r = (phi_oracle_result *)calloc(1, sizeof(*r));
    uint64_t na = 0, nv = 0;
    for (uint32_t h = 0; h < H; ++h) { na += ws[h].hash.n; nv += ws[h].vtx.n; }
    uint64_t *hashes = (uint64_t *)malloc((na ? na : 1) * 8);
    int32_t *owner = owner_vtx_out ? (int32_t *)malloc((na ? na : 1) * 4) : NULL;
    int32_t *arank = (int32_t *)calloc(na ? na : 1, 4), *awalk = (int32_t *)malloc((na ? na : 1) * 4);
    uint64_t *aoff = (uint64_t *)malloc((na + 1) * 8);
    int32_t *avtx = (int32_t *)malloc((nv ? nv : 1) * 4);
    uint64_t *mpw = (uint64_t *)calloc(H ? H : 1, 8), *apw = (uint64_t *)calloc(H ? H : 1, 8);
    uint64_t a = 0, vo = 0;
    aoff[0] = 0;
    for (uint32_t h = 0; h < H; ++h) {
        for (size_t i = 0; i < ws[h].hash.n; ++i) {
            hashes[a] = ws[h].hash.a[i]; awalk[a] = (int32_t)h;
            if (owner) owner[a] = ws[h].own.a[i];
            for (uint64_t q = ws[h].voff.a[i]; q < ws[h].voff.a[i + 1]; ++q) avtx[vo++] = ws[h].vtx.a[q];
            aoff[++a] = vo;
        }
        mpw[h] = apw[h] = ws[h].hash.n;
        uint64_t len = 0;
        for (uint64_t s = g->walk_off[h]; s < g->walk_off[h + 1]; ++s) len += g->seg_off[g->walk_vtx[s] + 1] - g->seg_off[g->walk_vtx[s]];
        r->path_kmer_positions += kmer_positions(len, prm->k, prm->w);
        free(ws[h].hash.a); free(ws[h].voff.a); free(ws[h].vtx.a); free(ws[h].own.a);
    }
    free(ws);
    if (owner_vtx_out) *owner_vtx_out = owner;
    r->n_walks = H; r->n_anchors = na; r->n_anchor_vtx = nv;
    r->anchor_walk = awalk; r->anchor_vtx = avtx;
    to_compact(r, arank, aoff, 0);
    r->minimizers_per_walk = mpw; r->anchors_per_walk = apw;
    r->path_minimizers_emitted = na;
    *out = r; *hashes_out = hashes;
    return PHI_OK;
}

/* ----------------------------------------------------------- filter / order */
typedef struct {
    int32_t walk;
    uint64_t seq;       /* insertion order inside the rank: (walk asc, path order) */
    const int32_t *v; int nv;
    char *key;          /* "v0_v1_..._" — ILP_index.cpp:680-683 */
    int group;          /* filled after grouping */
} hit_t;

static int cmp_hit_key_then_seq(const void *a, const void *b)
{
    const hit_t *x = (const hit_t *)a, *y = (const hit_t *)b;
    int c = strcmp(x->key, y->key);     /* std::map<std::string> order */
    if (c) return c;
    return x->seq < y->seq ? -1 : x->seq > y->seq;
}
static int cmp_hit_walk_key_seq(const void *a, const void *b)
{
    const hit_t *x = (const hit_t *)a, *y = (const hit_t *)b;
    if (x->walk != y->walk) return x->walk < y->walk ? -1 : 1;
    return cmp_hit_key_then_seq(a, b);
}

typedef struct { uint64_t h; int32_t w; } hw_t;
static int cmp_hw(const void *a, const void *b)
{
    const hw_t *x = (const hw_t *)a, *y = (const hw_t *)b;
    if (x->h != y->h) return x->h < y->h ? -1 : 1;
    return x->w < y->w ? -1 : x->w > y->w;
}

int phi_oracle_index_run(const phi_graph_view *g, const phi_reads_view *rd, const phi_index_params *prm,
                         int n_threads, phi_oracle_result **out)
{
    if (!g || !rd || !out || !check_params(prm)) return PHI_ERR_ARG;
    const int k = prm->k, w = prm->w;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#endif
    phi_oracle_result *r = (phi_oracle_result *)calloc(1, sizeof(*r));

    /* ---- loop A: walk sketches, ILP_index.cpp:556-573 */
    phi_oracle_result *wsr = 0; uint64_t *whash = 0;
    int rc = phi_oracle_sketch_walks(g, prm, n_threads, &wsr, &whash);
    if (rc) { free(r); return rc; }

    /* ---- the -d1 statistic, ILP_index.cpp:565-606: distinct walk-minimizer hashes by the number of walks they occur in */
    if (prm->debug) {
        const uint64_t na = wsr->n_anchors;
        hw_t *hw = (hw_t *)malloc((na ? na : 1) * sizeof(hw_t));
        for (uint64_t a = 0; a < na; ++a) { hw[a].h = whash[a]; hw[a].w = wsr->anchor_walk[a]; }
        if (na) qsort(hw, na, sizeof(hw_t), cmp_hw);
        uint64_t *hist = (uint64_t *)calloc((size_t)g->n_walks + 1, 8);
        uint64_t distinct = 0;
        for (uint64_t i = 0; i < na;) {
            uint64_t j = i, walks = 0;
            while (j < na && hw[j].h == hw[i].h) { if (j == i || hw[j].w != hw[j - 1].w) ++walks; ++j; }
            hist[walks]++; ++distinct;
            i = j;
        }
        free(hw);
        r->n_walk_kmers = distinct; r->shared_kmer_hist = hist;
    }

    /* ---- loop B: read sketches + Sp_R, ILP_index.cpp:615-636 */
    uint64_t R = rd->n_reads;
    uint64_t **rh = (uint64_t **)calloc(R ? R : 1, sizeof(uint64_t *));
    int64_t *rn = (int64_t *)calloc(R ? R : 1, sizeof(int64_t));
    #pragma omp parallel for num_threads(n_threads) schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)R; ++i)
        rn[i] = phi_oracle_read_hashes(rd->read_bases + rd->read_off[i], rd->read_off[i + 1] - rd->read_off[i], k, w, &rh[i]);
    vec_u64 all = {0, 0, 0};
    for (uint64_t i = 0; i < R; ++i) {
        for (int64_t j = 0; j < rn[i]; ++j) push_u64(&all, rh[i][j]);   /* Sp_R[hash]++ :623-629 */
        r->read_minimizers_emitted += (uint64_t)rn[i];
        r->read_kmer_positions += kmer_positions(rd->read_off[i + 1] - rd->read_off[i], k, w);
        free(rh[i]);
    }
    free(rh); free(rn);
    if (all.n) qsort(all.a, all.n, 8, cmp_u64);                          /* std::map iteration order == ascending unsigned */
    size_t ns = 0;
    for (size_t i = 0; i < all.n; ++i) if (i == 0 || all.a[i] != all.a[i - 1]) all.a[ns++] = all.a[i];
    const uint64_t *spec = all.a;                                        /* rank = index :631-635 */
    r->count_sp_r = (int32_t)ns;

    /* ---- loop C: match, ILP_index.cpp:495-526, :643-655.  Hits in (walk, path) order, bucketed by rank. */
    uint64_t na = wsr->n_anchors;
    uint64_t *wsr_off = (uint64_t *)malloc((na + 1) * 8);                /* offsets of the walk minimizers' vertex lists */
    wsr_off[0] = 0;
    for (uint64_t a = 0; a < na; ++a) wsr_off[a + 1] = wsr_off[a] + wsr->anchor_len[a];
    int64_t *hit_rank = (int64_t *)malloc((na ? na : 1) * sizeof(int64_t));
    uint64_t *per_rank = (uint64_t *)calloc(ns + 1, 8);
    uint64_t nhits = 0;
    for (uint64_t a = 0; a < na; ++a) {
        uint64_t key = whash[a]; size_t lo = 0, hi = ns;
        while (lo < hi) { size_t mid = (lo + hi) / 2; if (spec[mid] < key) lo = mid + 1; else hi = mid; }
        if (lo < ns && spec[lo] == key) { hit_rank[a] = (int64_t)lo; per_rank[lo + 1]++; ++nhits; } else hit_rank[a] = -1;
    }
    for (size_t i = 0; i < ns; ++i) per_rank[i + 1] += per_rank[i];
    uint64_t *by_rank = (uint64_t *)malloc((nhits ? nhits : 1) * 8), *cursor = (uint64_t *)malloc((ns + 1) * 8);
    memcpy(cursor, per_rank, (ns + 1) * 8);
    for (uint64_t a = 0; a < na; ++a) if (hit_rank[a] >= 0) by_rank[cursor[hit_rank[a]]++] = a;   /* stable: keeps (walk, path) order */
    free(cursor);
    r->path_hits = nhits;

    /* ---- filter + reorder, ILP_index.cpp:670-716 */
    const uint32_t H = g->n_walks;
    vec_i32 o_rank = {0, 0, 0}, o_walk = {0, 0, 0}, o_vtx = {0, 0, 0}; vec_u64 o_off = {0, 0, 0};
    push_u64(&o_off, 0);
    uint64_t *apw = (uint64_t *)calloc(H ? H : 1, 8);
    int64_t n_filtered = 0;
    const float thr = prm->threshold * (float)H;                         /* threshold * num_walks, float (:698) */
    for (size_t rk = 0; rk < ns; ++rk) {
        uint64_t b = per_rank[rk], e = per_rank[rk + 1], n = e - b;
        if (n == 0) continue;                                            /* empty map: all_haps stays false, rank retained */
        hit_t *hs = (hit_t *)malloc(n * sizeof(hit_t));
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t a = by_rank[b + i];
            hs[i].walk = wsr->anchor_walk[a]; hs[i].seq = i;
            hs[i].v = wsr->anchor_vtx + wsr_off[a]; hs[i].nv = (int)wsr->anchor_len[a];
            hs[i].key = (char *)malloc((size_t)hs[i].nv * 12 + 1);
            char *q = hs[i].key;
            for (int j = 0; j < hs[i].nv; ++j) q += sprintf(q, "%d_", hs[i].v[j]);   /* to_string(v) + "_" */
            *q = 0;
        }
        qsort(hs, n, sizeof(hit_t), cmp_hit_key_then_seq);
        int all_haps = 0;
        for (uint64_t i = 0; i < n;) {                                   /* group sizes :686-690, test :698 */
            uint64_t j = i + 1;
            while (j < n && strcmp(hs[j].key, hs[i].key) == 0) ++j;
            if ((float)(int32_t)(j - i) >= thr) { all_haps = 1; break; }
            i = j;
        }
        if (all_haps) ++n_filtered;                                      /* :711 */
        else {                                                           /* :705-709 then flattened by (walk, j) */
            qsort(hs, n, sizeof(hit_t), cmp_hit_walk_key_seq);
            for (uint64_t i = 0; i < n; ++i) {
                push_i32(&o_rank, (int32_t)rk); push_i32(&o_walk, hs[i].walk);
                for (int j = 0; j < hs[i].nv; ++j) push_i32(&o_vtx, hs[i].v[j]);
                push_u64(&o_off, o_vtx.n);
                apw[hs[i].walk]++;
            }
        }
        for (uint64_t i = 0; i < n; ++i) free(hs[i].key);
        free(hs);
    }
    free(hit_rank); free(per_rank); free(by_rank);

    r->n_walks = H; r->n_filtered = n_filtered;
    r->n_anchors = o_rank.n; r->n_anchor_vtx = o_vtx.n;
    r->spectrum = ns ? all.a : (free(all.a), (uint64_t *)calloc(1, 8));
    r->anchor_walk = o_walk.a ? o_walk.a : (int32_t *)calloc(1, 4);
    r->anchor_vtx = o_vtx.a ? o_vtx.a : (int32_t *)calloc(1, 4);
    r->count_sp_r = (int32_t)ns;
    to_compact(r, o_rank.a ? o_rank.a : (int32_t *)calloc(1, 4), o_off.a, (int32_t)ns);
    free(wsr_off);
    uint64_t *mpw = (uint64_t *)calloc(H ? H : 1, 8);
    memcpy(mpw, wsr->minimizers_per_walk, (size_t)H * 8);
    r->minimizers_per_walk = mpw; r->anchors_per_walk = apw;
    r->path_kmer_positions = wsr->path_kmer_positions;
    r->path_minimizers_emitted = wsr->path_minimizers_emitted;
    phi_oracle_result_free(wsr); free(whash);
    *out = r;
    return PHI_OK;
}

void phi_oracle_result_free(phi_oracle_result *r)
{
    if (!r) return;
    free((void *)r->spectrum); free((void *)r->rank_off); free((void *)r->anchor_walk);
    free((void *)r->anchor_len); free((void *)r->anchor_vtx);
    free((void *)r->minimizers_per_walk); free((void *)r->anchors_per_walk); free((void *)r->shared_kmer_hist);
    free(r);
}
void phi_oracle_free(void *p) { free(p); }
