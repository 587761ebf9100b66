// Host-side ingest: GFA -> flat graph view, FASTA/FASTQ -> flat reads view, written from scratch.
// What the reference does with gfatools + read_gfa + kseq (serial text I/O into nested std::vector / std::string
// containers, /root/reference/src/gfa-io.cpp:462-508, src/ILP_index.cpp:20-155, :313-328, src/kseq.h:192-232) is done here
// straight into the buffers include/phi_gpu_index.h describes, so nothing has to be re-flattened before the upload.
//
// Semantics kept (SURVEY.md §9 rule 11):
//   * only S, L and W records are looked at (gfa-io.cpp:493-495); lines shorter than 3 bytes or without a tab in column 2
//     are skipped (:492); P-lines are ignored
//   * vertex id = segment index by first appearance on an S- or L-line (gfa_add_seg, gfa-base.cpp:75-96)
//   * segment sequence '*' = no sequence (length from LN:i: only; the reference keeps no bases either)
//   * W-line: sample, haplotype index, contig, start, end, walk; steps naming an unknown segment are dropped
//     (gfa-io.cpp:399-405); walk name = sample + "." + hap (ILP_index.cpp:98)
//   * walk flip (gfa-io.cpp:64-115): the first orientation a segment is seen in (over all walks, in order) is its reference
//     strand; a walk with more steps against than with the reference strands is reverse-complemented
//   * a walk that still has a reverse-strand step is an error (ILP_index.cpp:104-107)
//   * adjacency = the arcs leaving forward vertices after symmetrisation (every L-line also gives the reverse-complement
//     arc, gfa-base.cpp:421-430), strands dropped (ILP_index.cpp:77-86); Kahn order with a FIFO queue seeded in vertex
//     order (:116-147).  The order of a vertex's arcs (which only breaks ties between equally valid topological orders;
//     the reference's comes out of gfatools' in-place radix sort) is L-line order here: top_order_map may differ from the
//     reference's in such ties, the front end's results cannot (SURVEY.md §9 rule 8) as long as every walk step follows an
//     L-line: phi_host_graph_unlinked_steps() counts the steps that do not (0 for every graph a pangenome builder writes).
//   * reads: kseq_read — header at the next '>' or '@', name up to the first white space, sequence lines concatenated
//     with one trailing '\r' stripped, FASTQ quality skipped and length-checked; parsing stops at the first malformed
//     record (kseq returns -2 and read_ip_reads' loop ends)
#include "../../include/phi_gpu_index.h"

#include <zlib.h>
#include <time.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

// PHI_HOST_TIMES=1: phase times of the loaders on stderr
struct PhaseTimer {
    bool on; double t0;
    static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
    PhaseTimer() : on(getenv("PHI_HOST_TIMES") != nullptr), t0(now()) {}
    void lap(const char *what) { if (!on) return; const double t = now(); fprintf(stderr, "[phi_host] %-24s %.3f s\n", what, t - t0); t0 = t; }
};

bool slurp(const char *path, std::string &out, std::string &err)
{
    // size hint: the gzip trailer holds the uncompressed size mod 2^32 (a plain file: its own size); the text is inflated
    // straight into the string, no staging copy
    size_t hint = 1 << 20;
    if (FILE *raw = fopen(path, "rb")) {
        unsigned char magic[2] = {0, 0}, tail[4];
        if (fread(magic, 1, 2, raw) == 2 && fseek(raw, 0, SEEK_END) == 0) {
            const long fsz = ftell(raw);
            if (magic[0] == 0x1f && magic[1] == 0x8b) {
                if (fsz >= 18 && fseek(raw, -4, SEEK_END) == 0 && fread(tail, 1, 4, raw) == 4)
                    hint = (size_t)tail[0] | (size_t)tail[1] << 8 | (size_t)tail[2] << 16 | (size_t)tail[3] << 24;
                if (hint < (size_t)fsz) hint = (size_t)fsz * 4;                 // wrapped or multi-member: only a starting point
            } else if (fsz > 0) hint = (size_t)fsz;
        }
        fclose(raw);
    }
    gzFile fp = gzopen(path, "rb");                       // reads plain files too
    if (!fp) { err = std::string("cannot open ") + path; return false; }
    gzbuffer(fp, 1 << 20);
    size_t have = 0;
    out.resize(hint + 1);
    for (;;) {
        if (have == out.size()) out.resize(out.size() + out.size() / 2 + (1 << 20));
        const size_t room = std::min<size_t>(out.size() - have, (size_t)1 << 30);
        const int n = gzread(fp, &out[have], (unsigned)room);
        if (n < 0) { err = std::string("read error in ") + path; gzclose(fp); return false; }
        if (n == 0) break;
        have += (size_t)n;
    }
    gzclose(fp);
    out.resize(have);
    return true;
}

}  // namespace

struct phi_host_graph {
    phi_graph_view view;
    std::vector<uint64_t> seg_off, walk_off;
    std::string seg_bases;
    std::vector<uint32_t> walk_vtx;
    std::vector<int32_t> top_order_map;
    uint64_t n_unlinked_steps = 0;
    std::vector<std::string> walk_names, seg_names;
    uint64_t n_links = 0;
};

struct phi_host_reads {
    phi_reads_view view;
    std::vector<uint64_t> read_off;
    std::string read_bases;
    std::vector<std::string> names;
};

static void set_err(char *err, size_t errlen, const std::string &m)
{
    if (err && errlen) { snprintf(err, errlen, "%s", m.c_str()); }
}

extern "C" int phi_host_graph_load(const char *gfa_path, phi_host_graph **out, char *err, size_t errlen)
{
    if (!gfa_path || !out) return PHI_ERR_ARG;
    *out = nullptr;
    std::string text, e;
    PhaseTimer pt;
    if (!slurp(gfa_path, text, e)) { set_err(err, errlen, e); return PHI_ERR_ARG; }
    pt.lap("inflate");
    phi_host_graph *G = new phi_host_graph();
    // Names and sequences are views into `text` while parsing: no per-field std::string, one open-addressing table keyed by the
    // bytes of the name (the reference goes through a khash of strdup'ed names, gfa-base.cpp:75-96).
    struct View { const char *p; uint32_t n; };
    struct NameTable {
        std::vector<uint32_t> slot; std::vector<View> *names; size_t mask;
        static uint64_t hash(const char *p, uint32_t n)
        {
            uint64_t h = 0xCBF29CE484222325ull ^ n;
            while (n >= 8) { uint64_t w; memcpy(&w, p, 8); h = (h ^ w) * 0x9E3779B97F4A7C15ull; h ^= h >> 29; p += 8; n -= 8; }
            uint64_t w = 0; memcpy(&w, p, n); h = (h ^ w) * 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
            return h;
        }
        void grow()
        {
            const size_t cap = slot.empty() ? (1u << 16) : slot.size() * 2;
            slot.assign(cap, 0xFFFFFFFFu); mask = cap - 1;
            for (uint32_t id = 0; id < names->size(); ++id) {
                uint32_t num;
                if (numeric((*names)[id].p, (*names)[id].n, num)) continue;          // lives in `direct`
                size_t s = hash((*names)[id].p, (*names)[id].n) & mask;
                while (slot[s] != 0xFFFFFFFFu) s = (s + 1) & mask;
                slot[s] = id;
            }
        }
        // Names of the form <fixed prefix><decimal number> (minigraph's s1, s2, ...; plain numbers from vg / pggb) skip the hash
        // table: the number indexes `direct`.  The prefix is the one of the first such name; a number with a leading zero, more
        // than 8 digits or another prefix is an ordinary name (hash table), so equal strings always take the same route.
        std::vector<uint32_t> direct; const char *prefix = nullptr; uint32_t prefix_len = 0; bool have_prefix = false; size_t hashed = 0;
        bool numeric(const char *p, uint32_t n, uint32_t &num)
        {
            uint32_t i = 0;
            while (i < n && (p[i] < '0' || p[i] > '9')) ++i;
            const uint32_t nd = n - i;
            if (nd == 0 || nd > 8 || (p[i] == '0' && nd > 1)) return false;
            uint32_t v = 0;
            for (uint32_t q = i; q < n; ++q) { if (p[q] < '0' || p[q] > '9') return false; v = v * 10 + (uint32_t)(p[q] - '0'); }
            if (v >= (1u << 24)) return false;
            if (!have_prefix) { have_prefix = true; prefix = p; prefix_len = i; }
            else if (i != prefix_len || memcmp(p, prefix, i) != 0) return false;
            num = v;
            return true;
        }
        // id of the name, or 0xFFFFFFFF; with add: the name gets the next id
        uint32_t find(const char *p, uint32_t n, bool add)
        {
            uint32_t num;
            if (numeric(p, n, num)) {
                if (num < direct.size() && direct[num] != 0xFFFFFFFFu) return direct[num];
                if (!add) return 0xFFFFFFFFu;
                if (num >= direct.size()) direct.resize(std::max<size_t>((size_t)num + 1, direct.size() * 2), 0xFFFFFFFFu);
                View v; v.p = p; v.n = n;
                direct[num] = (uint32_t)names->size(); names->push_back(v);
                return direct[num];
            }
            if (slot.empty() || (hashed + 1) * 10 > slot.size() * 7) grow();
            size_t s = hash(p, n) & mask;
            for (;; s = (s + 1) & mask) {
                const uint32_t id = slot[s];
                if (id == 0xFFFFFFFFu) break;
                if ((*names)[id].n == n && memcmp((*names)[id].p, p, n) == 0) return id;
            }
            if (!add) return 0xFFFFFFFFu;
            View v; v.p = p; v.n = n;
            slot[s] = (uint32_t)names->size(); names->push_back(v); ++hashed;
            return slot[s];
        }
    };
    std::vector<View> seg_name;                                      // per segment, in order of first appearance
    std::vector<View> seqs;                                          // per segment (n == 0: no sequence)
    NameTable name2id; name2id.names = &seg_name; name2id.mask = 0;
    std::vector<std::pair<uint32_t, uint32_t>> arcs;                // oriented vertices (seg << 1 | reverse)
    struct Walk { std::string sample; int hap; std::vector<uint32_t> v; };
    std::vector<Walk> walks;
    auto add_seg = [&](const char *b, const char *e) -> uint32_t {
        const uint32_t id = name2id.find(b, (uint32_t)(e - b), true);
        if (id == seqs.size()) { View none; none.p = b; none.n = 0; seqs.push_back(none); }
        return id;
    };
    std::vector<std::pair<const char *, const char *>> f;            // tab-separated fields of the current line
    const char *p = text.data(), *tend = p + text.size();
    while (p < tend) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(tend - p));
        const char *lend = nl ? nl : tend;
        const char *next_line = nl ? nl + 1 : tend;
        if (lend - p > 1 && lend[-1] == '\r') --lend;                    // kstream strips one trailing CR
        if (lend - p >= 3 && p[1] == '\t' && (p[0] == 'S' || p[0] == 'L' || p[0] == 'W')) {
            f.clear();
            const size_t want = p[0] == 'S' ? 2 : p[0] == 'L' ? 4 : 6;   // the fields that are looked at (the walk is the 6th)
            for (const char *q = p + 2;;) {
                const char *t = (const char *)memchr(q, '\t', (size_t)(lend - q));
                f.emplace_back(q, t ? t : lend);
                if (!t || f.size() == want) break;
                q = t + 1;
            }
            if (p[0] == 'S' && f.size() >= 2) {                          // name, sequence ('*': none)
                const uint32_t id = add_seg(f[0].first, f[0].second);
                View sq; sq.p = f[1].first; sq.n = (uint32_t)(f[1].second - f[1].first);
                if (sq.n == 1 && sq.p[0] == '*') sq.n = 0;
                seqs[id] = sq;
            } else if (p[0] == 'L' && f.size() >= 4) {                   // from, orientation, to, orientation [, overlap]
                const char oa = f[1].second > f[1].first ? f[1].first[0] : 0, ob = f[3].second > f[3].first ? f[3].first[0] : 0;
                if ((oa == '+' || oa == '-') && (ob == '+' || ob == '-')) {   // the reference tests the first byte only
                    uint32_t v = add_seg(f[0].first, f[0].second) << 1 | (oa != '+' ? 1u : 0u);
                    uint32_t w = add_seg(f[2].first, f[2].second) << 1 | (ob != '+' ? 1u : 0u);
                    arcs.emplace_back(v, w);
                    ++G->n_links;
                }
            } else if (p[0] == 'W' && f.size() >= 6) {                   // sample, haplotype, contig, start, end, walk
                Walk wk;
                wk.sample.assign(f[0].first, f[0].second); wk.hap = atoi(std::string(f[1].first, f[1].second).c_str());
                const char *c = f[5].first, *send = f[5].second;
                wk.v.reserve((size_t)(send - c) / 4);
                // tokens as gfa_parse_W cuts them (gfa-io.cpp:395-408): a token runs from one '>' / '<' to the next, and the FIRST
                // token starts at the first byte of the field whatever that byte is (it takes the orientation marker's place:
                // the name is what follows it); names that are no segment are dropped
                while (c < send) {
                    const char *d = c + 1;
                    while (d < send && *d != '>' && *d != '<') ++d;
                    const uint32_t id = name2id.find(c + 1, (uint32_t)(d - c - 1), false);
                    if (id != 0xFFFFFFFFu) wk.v.push_back(id << 1 | (*c == '<' ? 1u : 0u));
                    c = d;
                }
                walks.push_back(std::move(wk));
            }
        }
        p = next_line;
    }
    pt.lap("parse lines");
    G->seg_names.reserve(seg_name.size());
    for (const View &v : seg_name) G->seg_names.emplace_back(v.p, v.n);
    pt.lap("segment names");
    const uint32_t V = (uint32_t)seqs.size();
    // ---- walk flip (gfa-io.cpp:64-115)
    {
        std::vector<int8_t> strand(V, 0);
        for (auto &wk : walks) for (uint32_t v : wk.v) if (!strand[v >> 1]) strand[v >> 1] = (v & 1) ? -1 : 1;
        for (auto &wk : walks) {
            size_t with = 0, against = 0;
            for (uint32_t v : wk.v) (((v & 1) ? -1 : 1) == strand[v >> 1] ? with : against)++;
            if (with >= against) continue;
            const size_t n = wk.v.size();
            for (size_t j = 0; j < n / 2; ++j) { uint32_t t = wk.v[j] ^ 1; wk.v[j] = wk.v[n - 1 - j] ^ 1; wk.v[n - 1 - j] = t; }
            if (n & 1) wk.v[n / 2] ^= 1;
        }
    }
    // ---- flat views
    G->seg_off.assign((size_t)V + 1, 0);
    for (uint32_t v = 0; v < V; ++v) G->seg_off[v + 1] = G->seg_off[v] + seqs[v].n;
    G->seg_bases.resize(G->seg_off[V]);
    for (uint32_t v = 0; v < V; ++v) if (seqs[v].n) memcpy(&G->seg_bases[G->seg_off[v]], seqs[v].p, seqs[v].n);
    G->walk_off.assign(1, 0);
    for (size_t h = 0; h < walks.size(); ++h) {
        for (uint32_t v : walks[h].v) {
            if (v & 1) {                                                       // ILP_index.cpp:104-107
                set_err(err, errlen, "Error: Walk " + std::to_string(h) + " has reverse strand vertices " + std::to_string(v));
                delete G;
                return PHI_ERR_UNSUPPORTED;
            }
            G->walk_vtx.push_back(v >> 1);
        }
        G->walk_off.push_back(G->walk_vtx.size());
        G->walk_names.push_back(walks[h].sample + "." + std::to_string(walks[h].hap));
    }
    pt.lap("walk flip + flat views");
    // ---- adjacency of the forward vertices after symmetrisation, Kahn order (ILP_index.cpp:77-154)
    {
        std::vector<std::vector<uint32_t>> adj(V);
        auto add = [&](uint32_t a, uint32_t b) {
            if (a & 1) return;                                                 // only arcs leaving a forward vertex count
            auto &l = adj[a >> 1];
            for (uint32_t x : l) if (x == b) return;                           // multi-arcs are removed by gfa_cleanup
            l.push_back(b);
        };
        for (auto &ab : arcs) { add(ab.first, ab.second); add(ab.second ^ 1, ab.first ^ 1); }
        std::vector<int32_t> indeg(V, 0);
        for (uint32_t v = 0; v < V; ++v) for (uint32_t w : adj[v]) indeg[w >> 1]++;
        std::queue<uint32_t> q;
        for (uint32_t v = 0; v < V; ++v) if (!indeg[v]) q.push(v);
        G->top_order_map.assign(V, 0);
        int32_t next = 0;
        while (!q.empty()) {
            uint32_t u = q.front(); q.pop();
            G->top_order_map[u] = next++;
            for (uint32_t w : adj[u]) if (--indeg[w >> 1] == 0) q.push(w >> 1);
        }
        // walk steps that no L-line backs: where there are none (the rule for real graphs) the order of an anchor's vertices is walk
        // order under ANY valid topological order, so the Kahn tie-breaks above cannot show; where there are some, the reference's own
        // tie-breaks (gfatools' arc sort) would decide the order of those two vertices inside an anchor, and ours may differ
        G->n_unlinked_steps = 0;
        for (size_t h = 0; h + 1 < G->walk_off.size(); ++h)
            for (uint64_t s = G->walk_off[h]; s + 1 < G->walk_off[h + 1]; ++s) {
                const uint32_t a = G->walk_vtx[s], b = G->walk_vtx[s + 1] << 1;
                bool linked = false;
                for (uint32_t x : adj[a]) linked |= x == b;
                G->n_unlinked_steps += linked ? 0 : 1;
            }
    }
    pt.lap("adjacency + Kahn order");
    G->view.n_vtx = V; G->view.seg_off = G->seg_off.data(); G->view.seg_bases = (const uint8_t *)G->seg_bases.data();
    G->view.n_walks = (uint32_t)walks.size(); G->view.walk_off = G->walk_off.data(); G->view.walk_vtx = G->walk_vtx.data();
    G->view.top_order_map = G->top_order_map.data();
    *out = G;
    return PHI_OK;
}

extern "C" const phi_graph_view *phi_host_graph_view(const phi_host_graph *g) { return g ? &g->view : nullptr; }
extern "C" const char *phi_host_graph_walk_name(const phi_host_graph *g, uint32_t h) { return g && h < g->walk_names.size() ? g->walk_names[h].c_str() : ""; }
extern "C" const char *phi_host_graph_segment_name(const phi_host_graph *g, uint32_t v) { return g && v < g->seg_names.size() ? g->seg_names[v].c_str() : ""; }
extern "C" uint64_t phi_host_graph_n_links(const phi_host_graph *g) { return g ? g->n_links : 0; }
extern "C" uint64_t phi_host_graph_unlinked_steps(const phi_host_graph *g) { return g ? g->n_unlinked_steps : 0; }
extern "C" void phi_host_graph_free(phi_host_graph *g) { delete g; }

extern "C" int phi_host_reads_load(const char *path, phi_host_reads **out, char *err, size_t errlen)
{
    if (!path || !out) return PHI_ERR_ARG;
    *out = nullptr;
    std::string text, e;
    if (!slurp(path, text, e)) { set_err(err, errlen, e); return PHI_ERR_ARG; }
    phi_host_reads *R = new phi_host_reads();
    R->read_off.assign(1, 0);
    const char *p = text.data(), *end = p + text.size();
    auto line_end = [&](const char *s) { const char *nl = (const char *)memchr(s, '\n', (size_t)(end - s)); return nl ? nl : end; };
    int last_char = 0;
    for (;;) {                                                                  // kseq_read (kseq.h:192-232)
        if (!last_char) { while (p < end && *p != '>' && *p != '@') ++p; if (p >= end) break; last_char = *p++; }
        if (p >= end) break;                                                    // header char at the very end: ks_getuntil returns -1
        const char *q = p;
        while (q < end && !isspace((unsigned char)*q)) ++q;                     // name
        std::string name(p, q);
        if (q < end && *q != '\n') q = line_end(q);                             // comment
        p = q < end ? q + 1 : end;
        std::string seq;
        int c = -1;
        while (p < end) {
            c = (unsigned char)*p++;
            if (c == '>' || c == '+' || c == '@') break;
            if (c == '\n') { c = -1; continue; }
            seq.push_back((char)c);
            const char *le = line_end(p);
            seq.append(p, le);
            p = le < end ? le + 1 : end;
            if (seq.size() > 1 && seq.back() == '\r') seq.pop_back();
            c = -1;
        }
        last_char = (c == '>' || c == '@') ? c : 0;
        if (c == '+') {                                                         // FASTQ: skip the '+' line, read >= |seq| quality bytes
            const char *le = line_end(p);
            if (le >= end) break;                                               // -2: no quality string
            p = le + 1;
            size_t ql = 0; bool got = false;
            while (p < end || !got) {
                if (p >= end) break;
                const char *l2 = line_end(p);
                size_t n = (size_t)(l2 - p);
                ql += n;
                if (ql > 1 && n && l2[-1] == '\r') --ql;
                p = l2 < end ? l2 + 1 : end;
                got = true;
                if (ql >= seq.size()) break;
            }
            last_char = 0;
            if (ql != seq.size()) break;                                        // -2: truncated quality: the reference stops here
        }
        R->names.push_back(name);
        R->read_bases += seq;
        R->read_off.push_back(R->read_bases.size());
    }
    R->view.n_reads = R->read_off.size() - 1; R->view.read_off = R->read_off.data(); R->view.read_bases = (const uint8_t *)R->read_bases.data();
    *out = R;
    return PHI_OK;
}

extern "C" const phi_reads_view *phi_host_reads_view(const phi_host_reads *r) { return r ? &r->view : nullptr; }
extern "C" const char *phi_host_reads_name(const phi_host_reads *r, uint64_t i) { return r && i < r->names.size() ? r->names[i].c_str() : ""; }
extern "C" void phi_host_reads_free(phi_host_reads *r) { delete r; }
