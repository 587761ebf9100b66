// TEST INFRASTRUCTURE (CPU): phi_b200/csrc/fast_inflate.h against zlib.  Texts of several kinds (random bytes, DNA, FASTQ-like, GFA-like,
// runs, empty, tiny) are compressed by zlib at every level and strategy (stored, fixed and dynamic blocks; long and short matches) and
// decoded by inflate_raw: same bytes, same consumed length.  Then the compressed streams are cut and damaged at random: the decoder must
// return (an error, or some output inside the buffer) and never touch memory outside its buffers — the test is built with ASan + UBSan.
#include "../phi_b200/csrc/fast_inflate.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

static std::vector<uint8_t> deflate_raw(const std::string &text, int level, int strategy, int mem_level)
{
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, mem_level, strategy) != Z_OK) { fprintf(stderr, "deflateInit2 failed\n"); exit(2); }
    std::vector<uint8_t> out(deflateBound(&zs, (uLong)text.size()) + 64);
    zs.next_in = (Bytef *)text.data(); zs.avail_in = (uInt)text.size();
    zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { fprintf(stderr, "deflate failed\n"); exit(2); }
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return out;
}

int main(int argc, char **argv)
{
    const int rounds = argc > 1 ? atoi(argv[1]) : 40;
    std::mt19937_64 rng(12345);
    long checked = 0, damaged = 0, damaged_rejected = 0;
    for (int round = 0; round < rounds; ++round) {
        for (int kind = 0; kind < 9; ++kind) {
            size_t n = kind == 6 ? 0 : kind == 7 ? 1 + rng() % 40 : 1 + rng() % 300000;
            std::string t(n, '\0');
            switch (kind) {
            case 0: for (auto &c : t) c = (char)rng(); break;                                        // incompressible
            case 1: for (auto &c : t) c = "ACGT"[rng() & 3]; break;                                  // DNA
            case 2: {                                                                                // FASTQ-like
                size_t i = 0;
                while (i < n) {
                    std::string rec = "@read" + std::to_string(rng() % 100000) + "\n", s(150, 'A');
                    for (auto &c : s) c = "ACGTN"[rng() % 5 == 0 ? 4 : rng() & 3];
                    rec += s + "\n+\n" + std::string(150, (char)('!' + rng() % 40)) + "\n";
                    for (size_t j = 0; j < rec.size() && i < n; ++j) t[i++] = rec[j];
                }
                break;
            }
            case 3: { size_t i = 0; unsigned v = 1; while (i < n) { std::string tok = ">s" + std::to_string(v); v += 1 + rng() % 3;
                      for (size_t j = 0; j < tok.size() && i < n; ++j) t[i++] = tok[j]; } break; }   // W-line-like
            case 4: { char c = 'A'; size_t i = 0; while (i < n) { size_t run = 1 + rng() % 3000; for (size_t j = 0; j < run && i < n; ++j) t[i++] = c; c = (char)('A' + rng() % 26); } break; }   // long runs (distance 1 matches)
            case 5: { std::string unit(1 + rng() % 40, 'x'); for (auto &c : unit) c = (char)('a' + rng() % 4); for (size_t i = 0; i < n; ++i) t[i] = unit[i % unit.size()]; break; }   // short period
            case 8: for (auto &c : t) { unsigned v = 0; while (v < 255 && (rng() & 3)) ++v; c = (char)v; } break;   // geometric symbol frequencies: code lengths up to 15 (second-level tables)
            default: for (auto &c : t) c = (char)('a' + rng() % 3); break;
            }
            static const int strategies[] = {Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED};
            for (int level = 0; level <= 9; level += (round % 3 == 0 ? 1 : 3)) {
                const int strategy = strategies[(round + level + kind) % 5], mem_level = 1 + (int)(rng() % 9);
                const std::vector<uint8_t> z = deflate_raw(t, level, strategy, mem_level);
                std::vector<uint8_t> zin(z);                              // exact-size heap copies: ASan sees any byte read or written outside
                std::vector<uint8_t> out(n);
                size_t out_len = 0, used = 0, blocks = 0;
                const int rc = phi_inflate::inflate_raw(zin.data(), zin.size(), out.data(), out.size(), &out_len, &used, [&](size_t) { ++blocks; });
                if (rc != 0 || out_len != n || used != z.size() || (n && memcmp(out.data(), t.data(), n) != 0) || !blocks) {
                    fprintf(stderr, "MISMATCH kind %d level %d strategy %d n %zu: rc %d out %zu used %zu of %zu\n", kind, level, strategy, n, rc, out_len, used, z.size());
                    return 1;
                }
                ++checked;
                // trailing bytes after the stream must not be consumed
                zin.push_back(0xAB); zin.push_back(0xCD); zin.insert(zin.end(), 16, 0xEE);
                if (phi_inflate::inflate_raw(zin.data(), zin.size(), out.data(), out.size(), &out_len, &used, [](size_t) {}) != 0 || used != z.size() || out_len != n) {
                    fprintf(stderr, "MISMATCH with trailing bytes: kind %d level %d\n", kind, level); return 1;
                }
                // a buffer that is too small is an error, not an overrun
                if (n) { std::vector<uint8_t> small(n - 1 - (size_t)(rng() % std::min<size_t>(n, 9)));
                         if (phi_inflate::inflate_raw(z.data(), z.size(), small.data(), small.size(), &out_len, &used, [](size_t) {}) == 0) { fprintf(stderr, "small buffer accepted\n"); return 1; } }
                // cut and damaged streams
                for (int d = 0; d < 6; ++d) {
                    std::vector<uint8_t> bad(z);
                    if (d < 2 && !bad.empty()) bad.resize(rng() % bad.size());
                    else for (int f = 0; f < 1 + d; ++f) if (!bad.empty()) bad[rng() % bad.size()] ^= (uint8_t)(1u << (rng() % 8));
                    std::vector<uint8_t> o2(n + (size_t)(rng() % 3));
                    const int rc2 = phi_inflate::inflate_raw(bad.data(), bad.size(), o2.data(), o2.size(), &out_len, &used, [](size_t) {});
                    ++damaged; damaged_rejected += rc2 != 0;
                    if (rc2 == 0 && (out_len > o2.size() || used > bad.size())) { fprintf(stderr, "lengths out of range on damaged input\n"); return 1; }
                }
            }
        }
    }
    printf("inflate_raw == zlib on %ld streams; %ld damaged streams handled (%ld rejected)\n", checked, damaged, damaged_rejected);
    return 0;
}
