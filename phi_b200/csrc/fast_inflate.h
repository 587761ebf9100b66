// Raw-deflate decoder for WHOLE buffers (RFC 1951), written for the host loaders: the compressed file is mapped, the text goes straight
// to its final place, and — unlike zlib's streaming inflate — nothing is ever suspended or resumed, so the hot loop keeps a 64-bit
// bit buffer in a register, resolves a literal / length / distance with one table look-up each (11-bit and 8-bit root tables with
// second-level tables for the rare longer codes) and copies matches eight bytes at a time.  It is only an accelerator: the caller
// verifies the gzip trailer (CRC-32 and size) of what comes out and falls back to zlib — which then decides what a damaged file
// means, the way the reference's gzread does — whenever this decoder reports an error or the check fails.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace phi_inflate {

enum { LIT_ROOT = 11, DIST_ROOT = 8, MAX_BITS = 15 };
enum Kind : uint32_t { K_LITERAL = 0, K_BASE = 1, K_END = 2, K_SUB = 3 };

// entry: bits 0-7 code bits to consume (0: invalid), 8-9 kind, 10-15 extra bits (K_BASE) or sub-table bits (K_SUB), 16-31 value
static inline uint32_t entry(uint32_t bits, uint32_t kind, uint32_t extra, uint32_t val) { return bits | kind << 8 | extra << 10 | val << 16; }

struct Table {
    uint32_t e[(1 << LIT_ROOT) + 1024];          // root table followed by the second-level tables (<= 2^(15-root) entries per long prefix, bounded below)
    uint32_t used;
};

// Canonical Huffman code -> decode table.  lens[i] = code length of symbol i (0: unused).  sym_entry(i, consumed_bits) makes the
// entry of symbol i.  Returns false for an over-subscribed code or when the second-level tables would not fit.
template <class F>
static bool build_table(const uint8_t *lens, int n, int root, Table &t, F sym_entry)
{
    uint16_t count[MAX_BITS + 1] = {0}, next[MAX_BITS + 2];
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    count[0] = 0;
    uint32_t left = 1;                            // Kraft check
    for (int l = 1; l <= MAX_BITS; ++l) { left <<= 1; if (count[l] > left) return false; left -= count[l]; }
    uint32_t code = 0;
    next[1] = 0;
    for (int l = 1; l <= MAX_BITS; ++l) { code = (code + count[l - 1]) << 1; next[l] = (uint16_t)code; }
    const uint32_t root_size = 1u << root;
    for (uint32_t i = 0; i < root_size; ++i) t.e[i] = 0;
    t.used = root_size;
    // codes in canonical order: by length, then by symbol.  Long codes that share their first `root` bits (reversed: their LOW root bits)
    // are consecutive in that order, so every second-level table is opened once, sized for the longest code of its prefix.
    uint16_t rev_code[320]; uint8_t ln[320]; uint16_t sym[320];
    int m = 0;
    {
        uint16_t nc[MAX_BITS + 2];
        memcpy(nc, next, sizeof nc);
        for (int l = 1; l <= MAX_BITS; ++l)
            for (int i = 0; i < n; ++i)
                if (lens[i] == l) {
                    uint32_t c = nc[l]++, r = 0;
                    for (int b = 0; b < l; ++b) r |= ((c >> b) & 1u) << (l - 1 - b);
                    rev_code[m] = (uint16_t)r; ln[m] = (uint8_t)l; sym[m] = (uint16_t)i; ++m;
                }
    }
    for (int j = 0; j < m; ++j) {
        const int l = ln[j];
        if (l <= root) {
            const uint32_t en = sym_entry(sym[j], (uint32_t)l);
            for (uint32_t i = rev_code[j]; i < root_size; i += 1u << l) t.e[i] = en;
        } else {
            const uint32_t prefix = rev_code[j] & (root_size - 1);
            if (!t.e[prefix]) {                  // open the second-level table of this prefix: bits = longest code with it - root
                int maxl = l;
                for (int q = j + 1; q < m; ++q) if ((rev_code[q] & (root_size - 1)) == prefix && ln[q] > maxl) maxl = ln[q];
                const uint32_t sub_bits = (uint32_t)(maxl - root), sub_size = 1u << sub_bits;
                if (t.used + sub_size > sizeof t.e / sizeof t.e[0]) return false;
                for (uint32_t i = 0; i < sub_size; ++i) t.e[t.used + i] = 0;
                t.e[prefix] = entry((uint32_t)root, K_SUB, sub_bits, t.used);
                t.used += sub_size;
            }
            const uint32_t p = t.e[prefix], sub_bits = (p >> 10) & 63, base = p >> 16;
            const uint32_t en = sym_entry(sym[j], (uint32_t)(l - root));
            for (uint32_t i = rev_code[j] >> root; i < (1u << sub_bits); i += 1u << (l - root)) t.e[base + i] = en;
        }
    }
    return true;
}

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

static inline uint32_t litlen_entry(uint32_t s, uint32_t bits)
{
    if (s < 256) return entry(bits, K_LITERAL, 0, s);
    if (s == 256) return entry(bits, K_END, 0, 0);
    if (s <= 285) return entry(bits, K_BASE, LEN_EXTRA[s - 257], LEN_BASE[s - 257]);
    return 0;                                     // 286, 287: never valid in a stream
}
static inline uint32_t dist_entry(uint32_t s, uint32_t bits) { return s < 30 ? entry(bits, K_BASE, DIST_EXTRA[s], DIST_BASE[s]) : 0; }

struct Reader {
    const uint8_t *in; size_t n, pos;             // next byte to load
    uint64_t buf; uint32_t cnt;                   // cnt valid bits in buf (low bits first)
    inline void refill()
    {
        if (pos + 8 <= n) {                       // branch-free bulk refill: top up to >= 56 bits
            uint64_t w; memcpy(&w, in + pos, 8);
            buf |= w << cnt;
            const uint32_t take = (63 - cnt) >> 3;
            pos += take; cnt += take << 3;
        } else {
            while (cnt <= 56 && pos < n) { buf |= (uint64_t)in[pos++] << cnt; cnt += 8; }
        }
    }
    inline uint32_t peek(uint32_t b) const { return (uint32_t)(buf & ((1ull << b) - 1)); }
    inline void drop(uint32_t b) { buf >>= b; cnt -= b; }
};

// Decodes one raw deflate stream from in[0, n) into out[0, cap).  Returns 0 when the final block ended cleanly (*out_len bytes written,
// *in_used bytes consumed), a non-zero code otherwise (bad data, output would exceed cap, input exhausted).  progress(pos), when
// given, is called after every block with the number of output bytes that are final.
template <class P>
static int inflate_raw(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len, size_t *in_used, P progress)
{
    Reader r = {in, n, 0, 0, 0};
    size_t op = 0;
    static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    Table *lit = new Table, *dist = new Table, *cl = new Table;
    struct Free { Table *a, *b, *c; ~Free() { delete a; delete b; delete c; } } guard = {lit, dist, cl};
    (void)guard;
    int last;
    do {
        r.refill();
        if (r.cnt < 3) return 1;
        last = (int)r.peek(1); r.drop(1);
        const uint32_t type = r.peek(2); r.drop(2);
        if (type == 0) {                          // stored
            r.drop(r.cnt & 7);                    // to the byte boundary
            r.refill();
            if (r.cnt < 32) return 1;
            const uint32_t len = r.peek(16); r.drop(16);
            const uint32_t nlen = r.peek(16); r.drop(16);
            if ((len ^ 0xFFFFu) != nlen) return 2;
            // the bytes still in the bit buffer come first
            size_t back = r.cnt >> 3;
            r.pos -= back; r.buf = 0; r.cnt = 0;
            if (r.pos + len > n) return 1;
            if (op + len > cap) return 3;
            if (len) memcpy(out + op, in + r.pos, len);
            op += len; r.pos += len;
        } else if (type == 1 || type == 2) {
            uint8_t lens[320];
            int hlit, hdist;
            if (type == 1) {
                hlit = 288; hdist = 30;
                for (int i = 0; i < 144; ++i) lens[i] = 8;
                for (int i = 144; i < 256; ++i) lens[i] = 9;
                for (int i = 256; i < 280; ++i) lens[i] = 7;
                for (int i = 280; i < 288; ++i) lens[i] = 8;
                for (int i = 0; i < 30; ++i) lens[288 + i] = 5;
            } else {
                r.refill();
                if (r.cnt < 14) return 1;
                hlit = (int)r.peek(5) + 257; r.drop(5);
                hdist = (int)r.peek(5) + 1; r.drop(5);
                const int hclen = (int)r.peek(4) + 4; r.drop(4);
                if (hlit > 286 || hdist > 30) return 2;
                uint8_t cll[19] = {0};
                for (int i = 0; i < hclen; ++i) { r.refill(); if (r.cnt < 3) return 1; cll[ORDER[i]] = (uint8_t)r.peek(3); r.drop(3); }
                if (!build_table(cll, 19, 7, *cl, [](uint32_t s, uint32_t bits) { return entry(bits, K_LITERAL, 0, s); })) return 2;
                int i = 0;
                while (i < hlit + hdist) {
                    r.refill();
                    const uint32_t e = cl->e[r.peek(7)];
                    const uint32_t bits = e & 255;
                    if (!bits || bits > r.cnt) return bits ? 1 : 2;
                    r.drop(bits);
                    const uint32_t s = e >> 16;
                    if (s < 16) { lens[i++] = (uint8_t)s; continue; }
                    uint32_t rep, val = 0;
                    if (s == 16) { if (!i) return 2; if (r.cnt < 2) return 1; val = lens[i - 1]; rep = 3 + r.peek(2); r.drop(2); }
                    else if (s == 17) { if (r.cnt < 3) return 1; rep = 3 + r.peek(3); r.drop(3); }
                    else { if (r.cnt < 7) return 1; rep = 11 + r.peek(7); r.drop(7); }
                    if (i + (int)rep > hlit + hdist) return 2;
                    while (rep--) lens[i++] = (uint8_t)val;
                }
                if (!lens[256]) return 2;         // no end-of-block code
                // the distance lengths follow the literal/length lengths directly: move them to their own place
                memmove(lens + 288, lens + hlit, (size_t)hdist);
                for (int q = hlit; q < 288; ++q) lens[q] = 0;
            }
            if (!build_table(lens, type == 1 ? 288 : hlit, LIT_ROOT, *lit, litlen_entry)) return 2;
            if (!build_table(lens + 288, hdist, DIST_ROOT, *dist, dist_entry)) return 2;
            for (;;) {
                if (r.pos + 8 <= n && op + 336 <= cap) {
                    // fast iteration: the refill leaves >= 56 bits (a length + distance pair takes <= 48) and there is room for a run of
                    // literals (<= 63: one per bit) or the longest match plus the eight-byte copy slack, so nothing below checks input
                    // or output bounds
                    r.refill();
                    uint32_t e = lit->e[r.peek(LIT_ROOT)];
                    if (!(e & 0x300)) {                                  // a run of literals from one refill: while a whole code (<= 15 bits) is there
                        if (!(e & 255)) return 2;
                        bool more = false;
                        for (;;) {
                            r.drop(e & 255); out[op++] = (uint8_t)(e >> 16);
                            if (r.cnt < MAX_BITS) break;
                            e = lit->e[r.peek(LIT_ROOT)];
                            if ((e & 0x300) || !(e & 255)) { more = r.cnt >= 48 && (e & 255); break; }
                        }
                        if (!more) continue;                             // (else: the length / end code that ended the run is decoded from the same bits)
                    }
                    if (((e >> 8) & 3) == K_SUB) { r.drop(LIT_ROOT); e = lit->e[(e >> 16) + r.peek((e >> 10) & 63)]; }
                    if (!(e & 255)) return 2;
                    r.drop(e & 255);
                    const uint32_t kind = (e >> 8) & 3;
                    if (kind == K_LITERAL) { out[op++] = (uint8_t)(e >> 16); continue; }
                    if (kind == K_END) break;
                    const uint32_t xl = (e >> 10) & 63;
                    const uint32_t len = (e >> 16) + r.peek(xl);
                    r.drop(xl);
                    uint32_t d = dist->e[r.peek(DIST_ROOT)];
                    if (((d >> 8) & 3) == K_SUB) { r.drop(DIST_ROOT); d = dist->e[(d >> 16) + r.peek((d >> 10) & 63)]; }
                    if (!(d & 255)) return 2;
                    r.drop(d & 255);
                    const uint32_t xd = (d >> 10) & 63;
                    const size_t dd = (size_t)(d >> 16) + r.peek(xd);
                    r.drop(xd);
                    if (dd > op) return 2;
                    uint8_t *dst = out + op;
                    const uint8_t *src = dst - dd;
                    if (dd >= 8) {
                        for (uint32_t i = 0; i < len; i += 8) { uint64_t w; memcpy(&w, src + i, 8); memcpy(dst + i, &w, 8); }
                    } else if (dd == 1) {
                        memset(dst, src[0], len);
                    } else {
                        for (uint32_t i = 0; i < len; ++i) dst[i] = src[i];
                    }
                    op += len;
                    continue;
                }
                r.refill();
                uint32_t e = lit->e[r.peek(LIT_ROOT)];
                if (((e >> 8) & 3) == K_SUB) { if (r.cnt < LIT_ROOT) return 1; r.drop(LIT_ROOT); e = lit->e[(e >> 16) + r.peek((e >> 10) & 63)]; }
                uint32_t bits = e & 255;
                if (!bits) return 2;
                if (bits > r.cnt) return 1;
                r.drop(bits);
                const uint32_t kind = (e >> 8) & 3;
                if (kind == K_LITERAL) {
                    if (op >= cap) return 3;
                    out[op++] = (uint8_t)(e >> 16);
                    // a second literal from the same refill (>= 56 bits were there; two codes take <= 30)
                    e = lit->e[r.peek(LIT_ROOT)];
                    if (((e >> 8) & 3) != K_LITERAL || (e & 255) == 0 || (e & 255) > r.cnt || op >= cap) continue;
                    r.drop(e & 255);
                    out[op++] = (uint8_t)(e >> 16);
                    continue;
                }
                if (kind == K_END) break;
                const uint32_t xl = (e >> 10) & 63;
                if (xl > r.cnt) return 1;
                const uint32_t len = (e >> 16) + r.peek(xl);
                r.drop(xl);
                uint32_t d = dist->e[r.peek(DIST_ROOT)];
                if (((d >> 8) & 3) == K_SUB) { if (r.cnt < DIST_ROOT) return 1; r.drop(DIST_ROOT); d = dist->e[(d >> 16) + r.peek((d >> 10) & 63)]; }
                bits = d & 255;
                if (!bits) return 2;
                if (bits > r.cnt) return 1;
                r.drop(bits);
                const uint32_t xd = (d >> 10) & 63;
                if (xd > r.cnt) { r.refill(); if (xd > r.cnt) return 1; }
                const size_t dd = (size_t)(d >> 16) + r.peek(xd);
                r.drop(xd);
                if (dd > op) return 2;            // before the start of the output (no preset dictionary)
                if (op + len > cap) return 3;
                uint8_t *dst = out + op;
                const uint8_t *src = dst - dd;
                if (dd >= 8 && op + len + 8 <= cap) {       // eight bytes at a time (may write up to 7 bytes past the match: room checked)
                    for (uint32_t i = 0; i < len; i += 8) { uint64_t w; memcpy(&w, src + i, 8); memcpy(dst + i, &w, 8); }
                } else {
                    for (uint32_t i = 0; i < len; ++i) dst[i] = src[i];
                }
                op += len;
            }
        } else return 2;
        progress(op);
    } while (!last);
    // bytes loaded into the bit buffer but not used belong to what follows the stream
    *in_used = r.pos - (r.cnt >> 3);
    *out_len = op;
    return 0;
}

// gzip member header (RFC 1952) at b[0, n): offset of the deflate data, or 0 when this is not a plain gzip header.
static inline size_t gzip_header_size(const uint8_t *b, size_t n)
{
    if (n < 18 || b[0] != 0x1f || b[1] != 0x8b || b[2] != 8 || (b[3] & 0xE0)) return 0;
    const uint8_t flg = b[3];
    size_t o = 10;
    if (flg & 4) { if (o + 2 > n) return 0; o += 2 + ((size_t)b[o] | (size_t)b[o + 1] << 8); }
    if (flg & 8) { while (o < n && b[o]) ++o; ++o; }
    if (flg & 16) { while (o < n && b[o]) ++o; ++o; }
    if (flg & 2) o += 2;
    return o + 8 <= n ? o : 0;
}

}  // namespace phi_inflate
