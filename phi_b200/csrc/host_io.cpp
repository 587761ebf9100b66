// Host-side ingest: GFA -> flat graph view, FASTA/FASTQ -> flat reads view, written from scratch.
// What the reference does with gfatools + read_gfa + kseq (serial text I/O into nested std::vector / std::string
// containers, /root/reference/src/gfa-io.cpp:462-508, src/ILP_index.cpp:20-155, :313-328, src/kseq.h:192-232) is done here
// straight into the buffers include/phi_gpu_index.h describes, so nothing has to be re-flattened before the upload.
// Text arrives while it is parsed: single-member gzip through a reader thread, bgzip (BGZF) files through parallel inflate of the
// members (TextStream).
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
#include "fast_inflate.h"

#include <zlib.h>
#include <time.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
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

// f(i) for i in [0, n) on up to 16 host threads (items are handed out one by one: walks differ in length)
template <class F>
void parallel_for(size_t n, F f)
{
    const unsigned T = (unsigned)std::min<size_t>(std::min<size_t>(n, 16), std::max(1u, std::thread::hardware_concurrency()));
    if (T <= 1) { for (size_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t) th.emplace_back([&]() { for (;;) { const size_t i = next.fetch_add(1); if (i >= n) break; f(i); } });
    for (auto &x : th) x.join();
}

// A BGZF file (bgzip: what pangenome graphs, VCFs and many read sets are compressed with) is a chain of independent gzip members of
// at most 64 KB of text each, every member header carrying the member's compressed size ("BC" extra subfield) and every trailer its
// text size — so the members can be found by hopping from header to header and inflated in parallel straight to their final places.
struct BgzfIndex {
    struct Block { size_t in_off, in_len, out_off; uint32_t out_len, crc; };   // in_off / in_len: the raw deflate data of the member
    std::vector<Block> blocks;
    size_t total = 0;                                                           // text bytes
    const unsigned char *map = nullptr; size_t map_len = 0;                     // the compressed file (mmap)
    BgzfIndex() = default;
    BgzfIndex(const BgzfIndex &) = delete; BgzfIndex &operator=(const BgzfIndex &) = delete;
    BgzfIndex(BgzfIndex &&o) noexcept : blocks(std::move(o.blocks)), total(o.total), map(o.map), map_len(o.map_len) { o.map = nullptr; o.map_len = 0; }
    ~BgzfIndex() { if (map) munmap((void *)map, map_len); }
};

// true iff the WHOLE file is a well-formed chain of BGZF members (anything else — plain gzip, a cut file, trailing bytes — is left
// to the serial gzread path, which treats it the way the reference's reader does)
bool bgzf_index(const char *path, BgzfIndex &ix)
{
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 28) { close(fd); return false; }
    const size_t n = (size_t)st.st_size;
    void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return false;
    ix.map = (const unsigned char *)m; ix.map_len = n;
    const unsigned char *b = ix.map;
    auto le16 = [&](size_t o) { return (size_t)b[o] | (size_t)b[o + 1] << 8; };
    auto le32 = [&](size_t o) { return (uint32_t)b[o] | (uint32_t)b[o + 1] << 8 | (uint32_t)b[o + 2] << 16 | (uint32_t)b[o + 3] << 24; };
    size_t o = 0, out = 0;
    while (o < n) {
        if (n - o < 28 || b[o] != 0x1f || b[o + 1] != 0x8b || b[o + 2] != 8 || b[o + 3] != 4) return false;   // FLG: FEXTRA and nothing else
        const size_t xlen = le16(o + 10);
        if (o + 12 + xlen + 8 > n) return false;
        size_t bsize = 0;
        for (size_t f = o + 12; f + 4 <= o + 12 + xlen;) {
            const size_t slen = le16(f + 2);
            if (b[f] == 'B' && b[f + 1] == 'C' && slen == 2 && f + 6 <= o + 12 + xlen) bsize = le16(f + 4) + 1;
            f += 4 + slen;
        }
        if (!bsize || bsize < 12 + xlen + 8 || o + bsize > n) return false;
        BgzfIndex::Block k;
        k.in_off = o + 12 + xlen; k.in_len = bsize - 12 - xlen - 8;
        k.crc = le32(o + bsize - 8); k.out_len = le32(o + bsize - 4); k.out_off = out;
        ix.blocks.push_back(k);
        out += k.out_len; o += bsize;
    }
    ix.total = out;
    return true;
}

// Inflated text that becomes available while it is being parsed.  Plain gzip: a reader thread inflates (one deflate stream cannot be
// split) into a buffer of fixed capacity — the size the gzip trailer promises — and the parser, written against `more()` /
// `line_end()` instead of a fixed end pointer, follows right behind it; if the trailer lied (multi-member or > 4 GB streams) the
// buffer overflows, `overflow` is set and the caller falls back to slurp().  BGZF: up to 16 threads inflate the members in file
// order, each to its final place, and the text is released to the parser as the finished prefix grows.
struct TextStream {
    std::string owned;                                                   // complete text handed in by the caller
    char *base = nullptr; size_t cap = 0; bool heap = false;
    std::atomic<size_t> avail{0};
    std::atomic<bool> done{false};
    bool failed = false, overflow = false;
    std::mutex mu; std::condition_variable cv;
    std::thread th;
    BgzfIndex bz; std::vector<std::thread> workers; std::vector<uint8_t> block_done; size_t frontier = 0; std::atomic<size_t> next_block{0};
    std::atomic<bool> stop{false};

    void alloc(size_t n)
    {
        cap = n; heap = true;
        // no zero fill (every byte is written before it is released); big buffers on huge pages: one fault per 2 MB instead of per 4 KB
        void *q = nullptr;
        if (n >= ((size_t)4 << 20) && !getenv("PHI_HOST_NO_THP") && posix_memalign(&q, (size_t)2 << 20, n) == 0) { base = (char *)q; madvise(q, n, MADV_HUGEPAGE); }
        else base = (char *)malloc(n ? n : 1);
    }
    explicit TextStream(std::string &&whole) : owned(std::move(whole)) { base = &owned[0]; cap = owned.size(); avail = cap; done = true; }   // nothing to wait for
    TextStream(const char *path, size_t capacity)
    {
        alloc(capacity);
        if (!base) { failed = true; done = true; return; }
        th = std::thread([this, path]() {
            if (fast_single_member(path)) { finish(); return; }
            if (overflow) { finish(); return; }                          // the fast decoder gave up after text had been released: the caller repeats with zlib
            gzFile fp = gzopen(path, "rb");
            if (!fp) { failed = true; finish(); return; }
            gzbuffer(fp, 1 << 20);
            size_t have = 0;
            for (;;) {
                if (have == cap) {                                      // more text than promised?
                    char probe;
                    if (gzread(fp, &probe, 1) > 0) overflow = true;
                    break;
                }
                const int n = gzread(fp, base + have, (unsigned)std::min<size_t>(cap - have, (size_t)4 << 20));
                if (n < 0) { failed = true; break; }
                if (n == 0) break;
                have += (size_t)n;
                { std::lock_guard<std::mutex> lk(mu); avail = have; }
                cv.notify_one();
            }
            gzclose(fp);
            finish();
        });
    }
    // A file that is exactly ONE gzip member whose trailer promises `cap` bytes: decoded from the mapped file by the whole-buffer
    // decoder (fast_inflate.h), block by block straight to its final place, then checked against the trailer's CRC-32 (in parallel
    // pieces).  false + nothing released: not such a file (the zlib loop takes over); false + overflow: it looked like one but the
    // decoder or the check failed after text had been released — the caller discards what it parsed and lets zlib decide.
    bool fast_single_member(const char *path)
    {
        if (getenv("PHI_HOST_ZLIB_ONLY")) return false;
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 18) { close(fd); return false; }
        const size_t n = (size_t)st.st_size;
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        close(fd);
        if (m == MAP_FAILED) return false;
        const unsigned char *b = (const unsigned char *)m;
        struct Unmap { void *m; size_t n; ~Unmap() { munmap(m, n); } } unmap = {m, n};
        (void)unmap;
        const size_t off = phi_inflate::gzip_header_size(b, n);
        if (!off) return false;
        size_t out_len = 0, used = 0;
        const int rc = phi_inflate::inflate_raw(b + off, n - off - 8, (unsigned char *)base, cap, &out_len, &used, [this](size_t have) {
            if (have > avail.load(std::memory_order_relaxed)) { { std::lock_guard<std::mutex> lk(mu); avail = have; } cv.notify_one(); }
        });
        auto le32 = [&](size_t o) { return (uint32_t)b[o] | (uint32_t)b[o + 1] << 8 | (uint32_t)b[o + 2] << 16 | (uint32_t)b[o + 3] << 24; };
        bool ok = rc == 0 && off + used + 8 == n && out_len == cap && le32(n - 4) == (uint32_t)out_len;
        if (ok) {                                                        // CRC-32 of the text in parallel pieces, combined
            const size_t P = std::max<size_t>(1, std::min<size_t>(16, out_len >> 20));
            std::vector<uLong> part(P);
            parallel_for(P, [&](size_t i) {
                const size_t a = out_len * i / P, z = out_len * (i + 1) / P;
                uLong c = crc32(0L, Z_NULL, 0);
                for (size_t p = a; p < z;) { const size_t step = std::min<size_t>(z - p, (size_t)1 << 30); c = crc32(c, (const Bytef *)base + p, (uInt)step); p += step; }
                part[i] = c;
            });
            uLong c = part[0];
            for (size_t i = 1; i < P; ++i) c = crc32_combine(c, part[i], (z_off_t)(out_len * (i + 1) / P - out_len * i / P));
            ok = (uint32_t)c == le32(n - 8);
        }
        if (!ok) { if (avail.load() > 0) overflow = true; return false; }
        { std::lock_guard<std::mutex> lk(mu); avail = out_len; }
        return true;
    }
    explicit TextStream(BgzfIndex &&index) : bz(std::move(index))
    {
        alloc(bz.total);
        if (!base) { failed = true; done = true; return; }
        const size_t nb = bz.blocks.size();
        block_done.assign(nb, 0);
        if (!nb) { done = true; return; }
        unsigned T = (unsigned)std::min<size_t>(std::min<size_t>(nb, 16), std::max(1u, std::thread::hardware_concurrency()));
        if (const char *e = getenv("PHI_HOST_INFLATE_THREADS")) T = (unsigned)std::max(1, std::min(64, atoi(e)));
        const bool zlib_only = getenv("PHI_HOST_ZLIB_ONLY") != nullptr;
        for (unsigned t = 0; t < T; ++t) workers.emplace_back([this, nb, zlib_only]() {
            z_stream zs; memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) { fail_block(); return; }
            for (;;) {
                const size_t i = next_block.fetch_add(1);
                if (i >= nb || stop.load(std::memory_order_relaxed)) break;
                const BgzfIndex::Block &k = bz.blocks[i];
                // the whole-buffer decoder first (fast_inflate.h); zlib decides when it does not produce exactly what the trailer promises
                size_t got = 0, used = 0;
                bool ok = !zlib_only && phi_inflate::inflate_raw(bz.map + k.in_off, k.in_len, (unsigned char *)base + k.out_off, k.out_len, &got, &used, [](size_t) {}) == 0 &&
                          got == k.out_len && used == k.in_len &&
                          (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)(base + k.out_off), k.out_len) == k.crc;
                if (!ok && (ok = inflateReset(&zs) == Z_OK)) {
                    zs.next_in = const_cast<Bytef *>(bz.map + k.in_off); zs.avail_in = (uInt)k.in_len;
                    zs.next_out = (Bytef *)(base + k.out_off); zs.avail_out = k.out_len;
                    const int rc = inflate(&zs, Z_FINISH);
                    ok = rc == Z_STREAM_END && zs.avail_out == 0 && zs.avail_in == 0 &&
                         (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)(base + k.out_off), k.out_len) == k.crc;
                }
                if (!ok) { fail_block(); break; }
                bool fin = false;
                {
                    std::lock_guard<std::mutex> lk(mu);
                    block_done[i] = 1;
                    while (frontier < nb && block_done[frontier]) ++frontier;
                    avail = frontier < nb ? bz.blocks[frontier].out_off : bz.total;
                    if (frontier == nb) { done = true; fin = true; }
                }
                cv.notify_all();
                if (fin) break;
            }
            inflateEnd(&zs);
        });
    }
    void fail_block() { stop = true; { std::lock_guard<std::mutex> lk(mu); failed = true; done = true; } cv.notify_all(); }
    void join() { if (th.joinable()) th.join(); for (std::thread &w : workers) if (w.joinable()) w.join(); }
    ~TextStream() { join(); if (heap) free(base); }
    void finish() { { std::lock_guard<std::mutex> lk(mu); done = true; } cv.notify_one(); }
    const char *begin() const { return base; }
    // is there a byte at p?  (waits for the reader when p is at the current end)
    bool more(const char *p)
    {
        const size_t off = (size_t)(p - base);
        if (off < avail.load(std::memory_order_acquire)) return true;
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&]() { return off < avail.load() || done.load(); });
        return off < avail.load();
    }
    // the '\n' that ends the line s lies in, or the end of the text
    const char *line_end(const char *s)
    {
        size_t searched = (size_t)(s - base);                            // no '\n' in [s, searched)
        for (;;) {
            const bool fin = done.load(std::memory_order_acquire);       // (read before avail: once done is seen, avail is final)
            const size_t a = avail.load(std::memory_order_acquire);
            if (searched < a) {
                const char *nl = (const char *)memchr(base + searched, '\n', a - searched);
                if (nl) return nl;
                searched = a;
            }
            if (fin) return base + a;
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&]() { return avail.load() > a || done.load(); });
        }
    }
    const char *end_now() { return base + avail.load(std::memory_order_acquire); }
};

// capacity for a TextStream: what slurp() uses as its size hint, exact for single-member gzip files below 4 GB and for plain files
size_t text_size_hint(const char *path)
{
    size_t hint = 0;
    if (FILE *raw = fopen(path, "rb")) {
        unsigned char magic[2] = {0, 0}, tail[4];
        if (fread(magic, 1, 2, raw) == 2 && fseek(raw, 0, SEEK_END) == 0) {
            const long fsz = ftell(raw);
            if (magic[0] == 0x1f && magic[1] == 0x8b) {
                if (fsz >= 18 && fseek(raw, -4, SEEK_END) == 0 && fread(tail, 1, 4, raw) == 4)
                    hint = (size_t)tail[0] | (size_t)tail[1] << 8 | (size_t)tail[2] << 16 | (size_t)tail[3] << 24;
                if (hint < (size_t)fsz) hint = 0;                        // wrapped or multi-member: no usable promise
            } else if (fsz > 0) hint = (size_t)fsz;
        }
        fclose(raw);
    }
    return hint;
}

}  // namespace

// A big array that is written completely before it is read: no zero fill (std::vector::resize would run one serial pass over
// gigabytes on chromosome-scale inputs), 2 MB aligned and marked for transparent huge pages.
template <class T> struct RawArray {
    T *p = nullptr; size_t n = 0;
    RawArray() = default;
    RawArray(const RawArray &) = delete; RawArray &operator=(const RawArray &) = delete;
    ~RawArray() { free(p); }
    bool alloc(size_t count)
    {
        free(p); p = nullptr; n = count;
        const size_t bytes = std::max<size_t>(count * sizeof(T), 1);
        void *q = nullptr;
        if (bytes >= ((size_t)4 << 20) && posix_memalign(&q, (size_t)2 << 20, bytes) == 0) { p = (T *)q; madvise(q, bytes, MADV_HUGEPAGE); }
        else p = (T *)malloc(bytes);
        return p != nullptr;
    }
    T *data() { return p; } const T *data() const { return p; }
    size_t size() const { return n; }
    T &operator[](size_t i) { return p[i]; } const T &operator[](size_t i) const { return p[i]; }
};

struct phi_host_graph {
    phi_graph_view view;
    std::vector<uint64_t> seg_off, walk_off;
    RawArray<char> seg_bases;
    RawArray<uint32_t> walk_vtx;
    std::vector<int32_t> top_order_map;
    uint64_t n_unlinked_steps = 0;
    std::vector<std::string> walk_names, seg_names;
    uint64_t n_links = 0;
};

struct phi_host_reads {
    phi_reads_view view;
    std::vector<uint64_t> read_off;
    std::string read_bases;
    std::string name_arena;              // the names back to back, each followed by a NUL
    std::vector<uint64_t> name_off;
    bool raw = false;                    // the parallel parse stitches its parts into these instead (no zero fill of gigabytes)
    RawArray<char> bases_raw, names_raw;
    const char *bases() const { return raw ? bases_raw.data() : read_bases.data(); }
    const char *names() const { return raw ? names_raw.data() : name_arena.c_str(); }
};

static void set_err(char *err, size_t errlen, const std::string &m)
{
    if (err && errlen) { snprintf(err, errlen, "%s", m.c_str()); }
}

extern "C" int phi_host_graph_load(const char *gfa_path, phi_host_graph **out, char *err, size_t errlen)
{
    if (!gfa_path || !out) return PHI_ERR_ARG;
    *out = nullptr;
    PhaseTimer pt;
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
        bool numeric(const char *p, uint32_t n, uint32_t &num, bool establish = true)
        {
            uint32_t i = 0;
            while (i < n && (p[i] < '0' || p[i] > '9')) ++i;
            const uint32_t nd = n - i;
            if (nd == 0 || nd > 8 || (p[i] == '0' && nd > 1)) return false;
            uint32_t v = 0;
            for (uint32_t q = i; q < n; ++q) { if (p[q] < '0' || p[q] > '9') return false; v = v * 10 + (uint32_t)(p[q] - '0'); }
            if (v >= (1u << 24)) return false;
            if (!have_prefix) { if (!establish) return false; have_prefix = true; prefix = p; prefix_len = i; }   // (a lookup never fixes the prefix: no name is in `direct` yet)
            else if (i != prefix_len || memcmp(p, prefix, i) != 0) return false;
            num = v;
            return true;
        }
        // id of the name, or 0xFFFFFFFF; with add: the name gets the next id
        // read-only (safe from several threads): id of the name if it is known, else 0xFFFFFFFF
        uint32_t lookup(const char *p, uint32_t n) const
        {
            uint32_t num;
            if (const_cast<NameTable *>(this)->numeric(p, n, num, false)) return num < direct.size() ? direct[num] : 0xFFFFFFFFu;
            if (slot.empty()) return 0xFFFFFFFFu;
            for (size_t s = hash(p, n) & mask;; s = (s + 1) & mask) {
                const uint32_t id = slot[s];
                if (id == 0xFFFFFFFFu) return id;
                if ((*names)[id].n == n && memcmp((*names)[id].p, p, n) == 0) return id;
            }
        }
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
    // a W-line is only cut into fields during the scan; its steps are looked up afterwards, all walks in parallel.  The reference drops
    // steps that name a segment not defined SO FAR (gfa-io.cpp:399-405): ids are handed out in order of first appearance, so "known
    // at that line" is "id < number of segments at that line".
    struct Walk { std::string sample; int hap; std::vector<uint32_t> v; const char *c, *send; uint32_t known; };
    std::vector<Walk> walks;
    auto add_seg = [&](const char *b, const char *e) -> uint32_t {
        const uint32_t id = name2id.find(b, (uint32_t)(e - b), true);
        if (id == seqs.size()) { View none; none.p = b; none.n = 0; seqs.push_back(none); }
        return id;
    };
    std::vector<std::pair<const char *, const char *>> f;            // tab-separated fields of the current line
    // The text arrives while it is scanned: a reader thread inflates into a buffer sized by the gzip trailer (TextStream); if that
    // promise does not hold the scan is repeated over the text inflated the plain way.  Names and sequences stay views into the text.
    std::unique_ptr<TextStream> text;
    auto scan = [&](TextStream &T) {
    seg_name.clear(); seqs.clear(); arcs.clear(); walks.clear(); G->n_links = 0;
    name2id = NameTable(); name2id.names = &seg_name; name2id.mask = 0;
    const char *p = T.begin();
    while (T.more(p)) {
        const char *lend = T.line_end(p);
        const char *next_line = T.more(lend) ? lend + 1 : lend;
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
                wk.c = f[5].first; wk.send = f[5].second; wk.known = (uint32_t)seg_name.size();
                walks.push_back(std::move(wk));
            }
        }
        p = next_line;
    }
    };
    bool scanned = false;
    {
        BgzfIndex bz;
        if (bgzf_index(gfa_path, bz)) {                                         // bgzip: members inflated in parallel, scanned as they land
            text.reset(new TextStream(std::move(bz)));
            scan(*text);
            text->join();
            scanned = !text->failed;
            pt.lap(scanned ? "inflate (bgzf) || scan" : "bgzf scan (discarded)");
        }
    }
    const size_t hint = scanned ? 0 : text_size_hint(gfa_path);
    if (hint) {
        text.reset(new TextStream(gfa_path, hint));
        scan(*text);
        text->join();
        if (text->failed && text->avail.load() == 0 && !text->overflow) { delete G; set_err(err, errlen, std::string("cannot open ") + gfa_path); return PHI_ERR_ARG; }
        scanned = !text->failed && !text->overflow;
        pt.lap(scanned ? "inflate || scan lines" : "streamed scan (discarded)");
    }
    if (!scanned) {
        std::string whole, e;
        if (!slurp(gfa_path, whole, e)) { delete G; set_err(err, errlen, e); return PHI_ERR_ARG; }
        pt.lap("inflate");
        text.reset(new TextStream(std::move(whole)));
        scan(*text);
        pt.lap("scan lines");
    }
    // tokens as gfa_parse_W cuts them (gfa-io.cpp:395-408): a token runs from one '>' / '<' to the next, and the FIRST token starts at
    // the first byte of the field whatever that byte is (it takes the orientation marker's place: the name is what follows it);
    // names that are no segment (yet, at that line) are dropped
    parallel_for(walks.size(), [&](size_t h) {
        Walk &wk = walks[h];
        const char *c = wk.c, *send = wk.send;
        wk.v.reserve((size_t)(send - c) / 4);
        while (c < send) {
            const char *d = c + 1;
            while (d < send && *d != '>' && *d != '<') ++d;
            const uint32_t id = name2id.lookup(c + 1, (uint32_t)(d - c - 1));
            if (id < wk.known) wk.v.push_back(id << 1 | (*c == '<' ? 1u : 0u));
            c = d;
        }
    });
    pt.lap("walk steps (parallel)");
    G->seg_names.reserve(seg_name.size());
    for (const View &v : seg_name) G->seg_names.emplace_back(v.p, v.n);
    pt.lap("segment names");
    const uint32_t V = (uint32_t)seqs.size();
    // ---- walk flip (gfa-io.cpp:64-115)
    // (no step of any walk is spelled '<' in the graphs pangenome builders write: then every segment's first orientation is forward,
    // nothing is against it and nothing flips — found out per walk in parallel, and the two serial passes below are skipped)
    std::vector<uint8_t> has_rev(walks.size(), 0);
    parallel_for(walks.size(), [&](size_t h) { uint32_t any = 0; for (uint32_t v : walks[h].v) any |= v; has_rev[h] = (uint8_t)(any & 1); });
    if (std::any_of(has_rev.begin(), has_rev.end(), [](uint8_t x) { return x != 0; })) {
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
    if (!G->seg_bases.alloc(G->seg_off[V])) { set_err(err, errlen, "out of memory"); delete G; return PHI_ERR_NOMEM; }
    for (uint32_t v = 0; v < V; ++v) if (seqs[v].n) memcpy(&G->seg_bases[G->seg_off[v]], seqs[v].p, seqs[v].n);
    G->walk_off.assign(1, 0);
    for (size_t h = 0; h < walks.size(); ++h) {
        G->walk_off.push_back(G->walk_off.back() + walks[h].v.size());
        G->walk_names.push_back(walks[h].sample + "." + std::to_string(walks[h].hap));
    }
    if (!G->walk_vtx.alloc(G->walk_off.back())) { set_err(err, errlen, "out of memory"); delete G; return PHI_ERR_NOMEM; }
    std::vector<int64_t> bad(walks.size(), -1);                                // first reverse-strand step of every walk
    parallel_for(walks.size(), [&](size_t h) {
        uint32_t *dst = G->walk_vtx.data() + G->walk_off[h];
        const std::vector<uint32_t> &v = walks[h].v;
        for (size_t j = 0; j < v.size(); ++j) { if ((v[j] & 1) && bad[h] < 0) bad[h] = (int64_t)v[j]; dst[j] = v[j] >> 1; }
    });
    for (size_t h = 0; h < walks.size(); ++h)
        if (bad[h] >= 0) {                                                     // ILP_index.cpp:104-107
            set_err(err, errlen, "Error: Walk " + std::to_string(h) + " has reverse strand vertices " + std::to_string(bad[h]));
            delete G;
            return PHI_ERR_UNSUPPORTED;
        }
    pt.lap("walk flip + flat views");
    // ---- adjacency of the forward vertices after symmetrisation, Kahn order (ILP_index.cpp:77-154)
    {
        // CSR over the forward source vertices, a vertex's arcs in L-line order (counting sort by source: stable), multi-arcs removed
        // (gfa_cleanup) by a scan of the vertex's own short list
        std::vector<uint64_t> adj_off((size_t)V + 1, 0);
        auto each_arc = [&](auto &&fn) {
            for (auto &ab : arcs) {
                if (!(ab.first & 1)) fn(ab.first >> 1, ab.second);              // only arcs leaving a forward vertex count
                if (!((ab.second ^ 1) & 1)) fn((ab.second ^ 1) >> 1, ab.first ^ 1);  // the reverse-complement arc of the same L-line
            }
        };
        each_arc([&](uint32_t a, uint32_t) { adj_off[a + 1]++; });
        for (uint32_t v = 0; v < V; ++v) adj_off[v + 1] += adj_off[v];
        std::vector<uint32_t> adj_to(adj_off[V]);
        std::vector<uint64_t> adj_end(adj_off.begin(), adj_off.end() - 1);      // v's list is adj_to[adj_off[v] .. adj_end[v])
        each_arc([&](uint32_t a, uint32_t b) {
            for (uint64_t i = adj_off[a]; i < adj_end[a]; ++i) if (adj_to[i] == b) return;
            adj_to[adj_end[a]++] = b;
        });
        struct Adj { const std::vector<uint64_t> &b, &e; const std::vector<uint32_t> &to; };
        const Adj adj = {adj_off, adj_end, adj_to};
        std::vector<int32_t> indeg(V, 0);
        for (uint32_t v = 0; v < V; ++v) for (uint64_t i = adj.b[v]; i < adj.e[v]; ++i) indeg[adj.to[i] >> 1]++;
        std::vector<uint32_t> q; q.reserve(V);                                  // FIFO: the order vertices enter is the order they leave
        for (uint32_t v = 0; v < V; ++v) if (!indeg[v]) q.push_back(v);
        G->top_order_map.assign(V, 0);
        int32_t next = 0;
        for (size_t head = 0; head < q.size(); ++head) {
            const uint32_t u = q[head];
            G->top_order_map[u] = next++;
            for (uint64_t i = adj.b[u]; i < adj.e[u]; ++i) if (--indeg[adj.to[i] >> 1] == 0) q.push_back(adj.to[i] >> 1);
        }
        // walk steps that no L-line backs: where there are none (the rule for real graphs) the order of an anchor's vertices is walk
        // order under ANY valid topological order, so the Kahn tie-breaks above cannot show; where there are some, the reference's own
        // tie-breaks (gfatools' arc sort) would decide the order of those two vertices inside an anchor, and ours may differ
        const size_t NWALK = G->walk_off.size() - 1;
        std::vector<uint64_t> unlinked(NWALK, 0);
        parallel_for(NWALK, [&](size_t h) {
            uint64_t n = 0;
            for (uint64_t s = G->walk_off[h]; s + 1 < G->walk_off[h + 1]; ++s) {
                const uint32_t a = G->walk_vtx[s], b = G->walk_vtx[s + 1] << 1;
                bool linked = false;
                for (uint64_t i = adj.b[a]; i < adj.e[a]; ++i) linked |= adj.to[i] == b;
                n += linked ? 0 : 1;
            }
            unlinked[h] = n;
        });
        G->n_unlinked_steps = 0;
        for (uint64_t n : unlinked) G->n_unlinked_steps += n;
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

// An uncompressed file, mapped (false: not there, empty, or it starts with the gzip magic).
struct PlainFile {
    const unsigned char *map = nullptr; size_t len = 0;
    bool open(const char *path)
    {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size <= 0) { close(fd); return false; }
        void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        close(fd);
        if (m == MAP_FAILED) return false;
        map = (const unsigned char *)m; len = (size_t)st.st_size;
        if (len >= 2 && map[0] == 0x1f && map[1] == 0x8b) { munmap(m, len); map = nullptr; len = 0; return false; }
        return true;
    }
    ~PlainFile() { if (map) munmap((void *)map, len); }
};

// A complete text in memory behind the interface the record parser is written against (TextStream: text that is still arriving).
struct FixedText {
    const char *b, *e;
    const char *begin() const { return b; }
    bool more(const char *p) const { return p < e; }
    const char *line_end(const char *s) const { const char *nl = (const char *)memchr(s, '\n', (size_t)(e - s)); return nl ? nl : e; }
};

// What a parse produces; phi_host_reads holds one, the parallel parse one per chunk of the text.
struct ReadSink {
    std::vector<uint64_t> read_off;      // cumulative bases, starts with 0
    std::string read_bases, name_arena;
    std::vector<uint64_t> name_off;
};

// kseq_read in a loop (kseq.h:192-232) over a text that may still be arriving: the parser never looks past what `more` grants.
// The loop's state between two records is a position and `last_char` (0, or the header byte '>' / '@' that the sequence loop of the
// previous record already consumed).  A parse can start in the middle of the text from such a state (hp: the header byte of its
// first record; NULL: at `from`, looking for one) and stops BEFORE the first record whose header byte lies at or behind `limit`
// (NULL: no limit), returning that header byte's position — NULL when the text, or a malformed record (-2: the reference's loop
// ends there), ended the parse.
template <class Text>
static const char *parse_reads_range(Text &T, const char *from, const char *hp, const char *limit, ReadSink *R, bool *malformed)
{
    if (R->read_off.empty()) R->read_off.assign(1, 0);
    const char *p = hp ? hp + 1 : from;
    int last_char = hp ? (unsigned char)*hp : 0;
    if (malformed) *malformed = false;
    for (;;) {
        if (!last_char) { while (T.more(p) && *p != '>' && *p != '@') ++p; if (!T.more(p)) break; last_char = *p++; }
        if (limit && p - 1 >= limit) return p - 1;
        if (!T.more(p)) break;                                                  // header char at the very end: ks_getuntil returns -1
        const char *q = p;
        while (T.more(q) && !isspace((unsigned char)*q)) ++q;                   // name
        const size_t name_at = R->name_arena.size();
        R->name_arena.append(p, q); R->name_arena.push_back('\0');
        if (T.more(q) && *q != '\n') q = T.line_end(q);                         // comment
        p = T.more(q) ? q + 1 : q;
        const size_t seq_at = R->read_bases.size();                             // the sequence goes straight to its final place
        int c = -1;
        while (T.more(p)) {
            c = (unsigned char)*p++;
            if (c == '>' || c == '+' || c == '@') break;
            if (c == '\n') { c = -1; continue; }
            R->read_bases.push_back((char)c);
            const char *le = T.line_end(p);
            R->read_bases.append(p, le);
            p = T.more(le) ? le + 1 : le;
            if (R->read_bases.size() - seq_at > 1 && R->read_bases.back() == '\r') R->read_bases.pop_back();
            c = -1;
        }
        const size_t seq_len = R->read_bases.size() - seq_at;
        last_char = (c == '>' || c == '@') ? c : 0;
        bool keep = true;
        if (c == '+') {                                                         // FASTQ: skip the '+' line, read >= |seq| quality bytes
            const char *le = T.line_end(p);
            if (!T.more(le)) keep = false;                                      // -2: no quality string
            else {
                p = le + 1;
                size_t ql = 0; bool got = false;
                while (T.more(p) || !got) {
                    if (!T.more(p)) break;
                    const char *l2 = T.line_end(p);
                    size_t n = (size_t)(l2 - p);
                    ql += n;
                    if (ql > 1 && n && l2[-1] == '\r') --ql;
                    p = T.more(l2) ? l2 + 1 : l2;
                    got = true;
                    if (ql >= seq_len) break;
                }
                last_char = 0;
                if (ql != seq_len) keep = false;                                // -2: truncated quality: the reference stops here
            }
        }
        if (!keep) { R->read_bases.resize(seq_at); R->name_arena.resize(name_at); if (malformed) *malformed = true; break; }
        R->name_off.push_back(name_at);
        R->read_off.push_back(R->read_bases.size());
    }
    return nullptr;
}

static void sink_to_reads(ReadSink &S, phi_host_reads *R)
{
    R->read_off.swap(S.read_off); R->read_bases.swap(S.read_bases); R->name_arena.swap(S.name_arena); R->name_off.swap(S.name_off);
    if (R->read_off.empty()) R->read_off.assign(1, 0);
}

static void parse_reads(TextStream &T, phi_host_reads *R)
{
    ReadSink S;
    S.read_bases.swap(R->read_bases);                                           // (keeps the caller's reserve)
    parse_reads_range(T, T.begin(), nullptr, nullptr, &S, nullptr);
    sink_to_reads(S, R);
}

// The same over a complete text, cut into chunks that are parsed side by side.  Where a record starts cannot be told from the
// middle of a FASTQ file (quality lines may begin with '@' or '>'), so every chunk GUESSES its first header — a line that begins
// with '>' , or with '@' when the line after next begins with '+' — and parses from there; afterwards the chain is checked: the
// parse of chunk i-1 must have stopped exactly at the header chunk i started from, in which case the two states are identical and
// the concatenation is what the serial loop produces.  The first link that does not hold (a wrong guess; a malformed record, after
// which the reference reads nothing) ends the chain, and the rest of the text is parsed serially from the true state.
static void parse_reads_parallel(const char *b, const char *e, phi_host_reads *R)
{
    FixedText T = {b, e};
    size_t chunk = 0;
    if (const char *env = getenv("PHI_HOST_PARSE_CHUNK")) chunk = (size_t)strtoull(env, nullptr, 10);    // tests: chunk borders everywhere
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (!chunk) chunk = std::max<size_t>((size_t)(e - b) / (hw * 4) + 1, (size_t)1 << 20);
    std::vector<const char *> starts;                                           // starts[0] = b (no guess), then one guessed header per chunk that has one
    starts.push_back(b);
    for (const char *c = b + chunk; c < e; c += chunk) {
        const char *lim = c + chunk < e ? c + chunk : e;
        const char *l = (const char *)memchr(c, '\n', (size_t)(e - c));
        for (l = l ? l + 1 : e; l < lim; ) {
            const char *nl = (const char *)memchr(l, '\n', (size_t)(e - l));
            if (*l == '>') break;
            if (*l == '@' && nl) {
                const char *nl2 = (const char *)memchr(nl + 1, '\n', (size_t)(e - nl - 1));
                if (nl2 && nl2 + 1 < e && nl2[1] == '+') break;
            }
            l = nl ? nl + 1 : e;
        }
        if (l < lim && l > starts.back()) starts.push_back(l);
    }
    const size_t n = starts.size();
    if (n == 1) { ReadSink S; parse_reads_range(T, b, nullptr, nullptr, &S, nullptr); sink_to_reads(S, R); return; }
    std::vector<ReadSink> part(n);
    std::vector<const char *> stop(n, nullptr);
    parallel_for(n, [&](size_t i) {
        part[i].read_bases.reserve((size_t)((i + 1 < n ? starts[i + 1] : e) - starts[i]) / 2 + 64);
        stop[i] = parse_reads_range(T, starts[i], i ? starts[i] : nullptr, i + 1 < n ? starts[i + 1] : nullptr, &part[i], nullptr);
    });
    size_t good = 1;                                                            // parts [0, good) are what the serial loop produces
    while (good < n && stop[good - 1] == starts[good]) ++good;
    ReadSink tail;
    if (good < n && stop[good - 1]) parse_reads_range(T, nullptr, stop[good - 1], nullptr, &tail, nullptr);   // the true state, serially to the end
    // (stop == NULL: the text or a malformed record ended the serial loop inside part good-1: nothing follows)
    std::vector<ReadSink *> seq;
    for (size_t i = 0; i < good; ++i) seq.push_back(&part[i]);
    if (good < n) seq.push_back(&tail);
    std::vector<uint64_t> base_at(seq.size() + 1, 0), name_at(seq.size() + 1, 0), read_at(seq.size() + 1, 0);
    for (size_t i = 0; i < seq.size(); ++i) {
        if (seq[i]->read_off.empty()) seq[i]->read_off.assign(1, 0);
        base_at[i + 1] = base_at[i] + seq[i]->read_bases.size(); name_at[i + 1] = name_at[i] + seq[i]->name_arena.size();
        read_at[i + 1] = read_at[i] + (seq[i]->read_off.size() - 1);
    }
    R->raw = R->bases_raw.alloc(base_at.back()) && R->names_raw.alloc(name_at.back() + 1);
    if (!R->raw) { R->read_bases.resize(base_at.back()); R->name_arena.resize(name_at.back()); }
    else R->names_raw[name_at.back()] = 0;
    char *bases_out = R->raw ? R->bases_raw.data() : &R->read_bases[0], *names_out = R->raw ? R->names_raw.data() : &R->name_arena[0];
    R->read_off.resize(read_at.back() + 1); R->name_off.resize(read_at.back());
    R->read_off[0] = 0;
    parallel_for(seq.size(), [&](size_t i) {
        const ReadSink &S = *seq[i];
        if (!S.read_bases.empty()) memcpy(bases_out + base_at[i], S.read_bases.data(), S.read_bases.size());
        if (!S.name_arena.empty()) memcpy(names_out + name_at[i], S.name_arena.data(), S.name_arena.size());
        for (size_t r = 0; r + 1 < S.read_off.size(); ++r) {
            R->read_off[read_at[i] + r + 1] = base_at[i] + S.read_off[r + 1];
            R->name_off[read_at[i] + r] = name_at[i] + S.name_off[r];
        }
    });
}

extern "C" int phi_host_reads_load(const char *path, phi_host_reads **out, char *err, size_t errlen)
{
    if (!path || !out) return PHI_ERR_ARG;
    *out = nullptr;
    PhaseTimer pt;
    phi_host_reads *R = new phi_host_reads();
    bool parsed = false;
    {
        BgzfIndex bz;
        if (bgzf_index(path, bz)) {                                             // bgzip: members inflated in parallel ...
            TextStream T(std::move(bz));
            if (T.cap >= ((size_t)8 << 20) || getenv("PHI_HOST_PARSE_CHUNK")) {   // ... then parsed in parallel chunks,
                T.join();
                pt.lap("inflate (bgzf)");
                if (!T.failed) { parse_reads_parallel(T.base, T.base + T.cap, R); pt.lap("parse (parallel chunks)"); }
            } else {                                                            // or, small files, parsed as they land
                R->read_bases.reserve(T.cap / 2 + 1024);
                parse_reads(T, R);
                T.join();
                pt.lap("inflate (bgzf) || parse");
            }
            parsed = !T.failed;
            if (!parsed) { delete R; R = new phi_host_reads(); }
        }
    }
    if (!parsed) {                                                              // not compressed: the file itself is the text
        PlainFile pf;
        if (pf.open(path)) {
            parse_reads_parallel((const char *)pf.map, (const char *)pf.map + pf.len, R);
            parsed = true;
            pt.lap("parse (mapped file, parallel chunks)");
        }
    }
    const size_t hint = parsed ? 0 : text_size_hint(path);
    if (hint) {                                                                 // streamed: the parser runs while the reader thread inflates
        TextStream T(path, hint);
        R->read_bases.reserve(hint / 2 + 1024);
        parse_reads(T, R);
        T.join();
        if (T.failed && T.avail.load() == 0 && !T.overflow) { delete R; set_err(err, errlen, std::string("cannot open ") + path); return PHI_ERR_ARG; }
        parsed = !T.failed && !T.overflow;
        pt.lap(parsed ? "inflate || parse" : "streamed pass (discarded)");
    }
    if (!parsed) {                                                              // no usable size promise: inflate everything, then parse
        delete R; R = new phi_host_reads();
        std::string text, e;
        if (!slurp(path, text, e)) { delete R; set_err(err, errlen, e); return PHI_ERR_ARG; }
        pt.lap("inflate");
        TextStream T(std::move(text));
        parse_reads(T, R);
        pt.lap("parse");
    }
    R->view.n_reads = R->read_off.size() - 1; R->view.read_off = R->read_off.data(); R->view.read_bases = (const uint8_t *)R->bases();
    *out = R;
    return PHI_OK;
}

extern "C" const phi_reads_view *phi_host_reads_view(const phi_host_reads *r) { return r ? &r->view : nullptr; }
extern "C" const char *phi_host_reads_name(const phi_host_reads *r, uint64_t i) { return r && i < r->name_off.size() ? r->names() + r->name_off[i] : ""; }
extern "C" void phi_host_reads_free(phi_host_reads *r) { delete r; }
