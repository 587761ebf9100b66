// phi_model_tables.hpp — the Gurobi-free parts of phi_model.hpp: integer-keyed open-addressing tables and the comparators that
// order the nodes of the expanded graph the way their names compare as std::strings.  tests/cpp/model_order_selftest.cpp checks
// them against std::map / std::string.
#ifndef PHI_MODEL_TABLES_HPP
#define PHI_MODEL_TABLES_HPP

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace phi_adapter {

// Names like "17_3_18_3" or "Flow_conservation_w_5_9" without one temporary std::string per piece: same text as the reference's
// std::to_string concatenations (decimal, '-' for negative numbers).
class Name {
public:
    Name() : n_(0) {}
    Name &s(const char *t) { while (*t) buf_[n_++] = *t++; return *this; }
    Name &c(char ch) { buf_[n_++] = ch; return *this; }
    Name &i(int64_t v)
    {
        char tmp[24]; int k = 0;
        uint64_t x = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
        do { tmp[k++] = (char)('0' + x % 10); x /= 10; } while (x);
        if (v < 0) buf_[n_++] = '-';
        while (k) buf_[n_++] = tmp[--k];
        return *this;
    }
    std::string str() const { return std::string(buf_, (size_t)n_); }
private:
    char buf_[160]; int n_;
};

// (u, v, j) -> index into a GRBVar pool; linear probing, grows by doubling
class EdgeVarTable {
public:
    EdgeVarTable() : mask_(0), used_(0) { rehash(1u << 16); }
    // returns the slot's pool index, or -1 after reserving the slot for `next_index`
    int64_t find_or_reserve(int32_t u, int32_t v, int32_t j, int64_t next_index)
    {
        if ((used_ + 1) * 10 > (mask_ + 1) * 7) rehash((mask_ + 1) * 2);
        size_t s = hash(u, v, j) & mask_;
        for (;; s = (s + 1) & mask_) {
            Slot &e = slots_[s];
            if (e.idx < 0) { e.u = u; e.v = v; e.j = j; e.idx = next_index; ++used_; return -1; }
            if (e.u == u && e.v == v && e.j == j) return e.idx;
        }
    }
    int64_t find(int32_t u, int32_t v, int32_t j) const
    {
        for (size_t s = hash(u, v, j) & mask_;; s = (s + 1) & mask_) {
            const Slot &e = slots_[s];
            if (e.idx < 0) return -1;
            if (e.u == u && e.v == v && e.j == j) return e.idx;
        }
    }
private:
    struct Slot { int32_t u, v, j; int64_t idx; };
    static size_t hash(int32_t u, int32_t v, int32_t j)
    {
        uint64_t x = ((uint64_t)(uint32_t)u << 32 | (uint32_t)v) * 0x9E3779B97F4A7C15ull ^ (uint64_t)(uint32_t)j * 0xD6E8FEB86659FD93ull;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        return (size_t)x;
    }
    void rehash(size_t n)
    {
        std::vector<Slot> old; old.swap(slots_);
        Slot empty; empty.u = empty.v = empty.j = 0; empty.idx = -1;
        slots_.assign(n, empty); mask_ = n - 1; used_ = 0;
        for (size_t i = 0; i < old.size(); ++i) if (old[i].idx >= 0) {
            size_t s = hash(old[i].u, old[i].v, old[i].j) & mask_;
            while (slots_[s].idx >= 0) s = (s + 1) & mask_;
            slots_[s] = old[i]; ++used_;
        }
    }
    std::vector<Slot> slots_; size_t mask_, used_;
};

// (u, i, v, j) -> index into a GRBVar pool (variables "u_i_v_j" between two walks); same scheme as EdgeVarTable
class CrossVarTable {
public:
    CrossVarTable() : mask_(0), used_(0) { rehash(1u << 16); }
    int64_t find_or_reserve(int32_t u, int32_t i, int32_t v, int32_t j, int64_t next_index)
    {
        if ((used_ + 1) * 10 > (mask_ + 1) * 7) rehash((mask_ + 1) * 2);
        const uint64_t k1 = (uint64_t)(uint32_t)u << 32 | (uint32_t)v, k2 = (uint64_t)(uint32_t)i << 32 | (uint32_t)j;
        for (size_t s = hash(k1, k2) & mask_;; s = (s + 1) & mask_) {
            Slot &e = slots_[s];
            if (e.idx < 0) { e.k1 = k1; e.k2 = k2; e.idx = next_index; ++used_; return -1; }
            if (e.k1 == k1 && e.k2 == k2) return e.idx;
        }
    }
private:
    struct Slot { uint64_t k1, k2; int64_t idx; };
    static size_t hash(uint64_t a, uint64_t b)
    {
        uint64_t x = a * 0x9E3779B97F4A7C15ull ^ (b + 0x7F4A7C15ull) * 0xD6E8FEB86659FD93ull;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        return (size_t)x;
    }
    void rehash(size_t n)
    {
        std::vector<Slot> old; old.swap(slots_);
        Slot empty; empty.k1 = empty.k2 = 0; empty.idx = -1;
        slots_.assign(n, empty); mask_ = n - 1; used_ = 0;
        for (size_t q = 0; q < old.size(); ++q) if (old[q].idx >= 0) {
            size_t s = hash(old[q].k1, old[q].k2) & mask_;
            while (slots_[s].idx >= 0) s = (s + 1) & mask_;
            slots_[s] = old[q]; ++used_;
        }
    }
    std::vector<Slot> slots_; size_t mask_, used_;
};

// key (a, b) -> dense index (assigned in order of first appearance)
class PairIndex {
public:
    PairIndex() : mask_(0), n_(0) { rehash(1u << 16); }
    uint32_t get(uint32_t a, uint32_t b, bool *is_new = 0)
    {
        if ((n_ + 1) * 10 > (mask_ + 1) * 7) rehash((mask_ + 1) * 2);
        const uint64_t key = (uint64_t)a << 32 | b;
        size_t s = mix(key) & mask_;
        for (;; s = (s + 1) & mask_) {
            if (slots_[s].idx == NONE) { slots_[s].key = key; slots_[s].idx = n_; if (is_new) *is_new = true; return n_++; }
            if (slots_[s].key == key) { if (is_new) *is_new = false; return slots_[s].idx; }
        }
    }
    bool find(uint32_t a, uint32_t b, uint32_t &idx) const
    {
        const uint64_t key = (uint64_t)a << 32 | b;
        for (size_t s = mix(key) & mask_;; s = (s + 1) & mask_) {
            if (slots_[s].idx == NONE) return false;
            if (slots_[s].key == key) { idx = slots_[s].idx; return true; }
        }
    }
    uint32_t size() const { return n_; }
private:
    enum { NONE = 0xFFFFFFFFu };
    struct Slot { uint64_t key; uint32_t idx; };
    static size_t mix(uint64_t x) { x *= 0x9E3779B97F4A7C15ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; return (size_t)x; }
    void rehash(size_t n)
    {
        std::vector<Slot> old; old.swap(slots_);
        Slot e; e.key = 0; e.idx = NONE;
        slots_.assign(n, e); mask_ = n - 1;
        for (size_t i = 0; i < old.size(); ++i) if (old[i].idx != NONE) {
            size_t s = mix(old[i].key) & mask_;
            while (slots_[s].idx != NONE) s = (s + 1) & mask_;
            slots_[s] = old[i];
        }
    }
    std::vector<Slot> slots_; size_t mask_; uint32_t n_;
};

inline int dec_digits(uint32_t v) { int d = 1; while (v >= 10) { v /= 10; ++d; } return d; }
// order of std::to_string(a) + tail_a vs std::to_string(b) + tail_b as strings, where both tails are empty (underscore == false: a
// proper prefix sorts first) or start with '_' (underscore == true: '_' sorts after every digit, so a proper prefix sorts last)
inline int dec_cmp(uint32_t a, uint32_t b, bool underscore)
{
    if (a == b) return 0;
    const int da = dec_digits(a), db = dec_digits(b);
    if (da == db) return a < b ? -1 : 1;
    uint32_t x = a, y = b;
    if (da < db) for (int i = 0; i < db - da; ++i) y /= 10; else for (int i = 0; i < da - db; ++i) x /= 10;
    if (x != y) return x < y ? -1 : 1;
    return ((da < db) != underscore) ? -1 : 1;          // the shorter one is a proper prefix of the longer one
}

struct XNode { uint32_t w, a, b; };                      // w == 0: A(v = a, walk = b), name "a_b";  w == 1: W(u = a, v = b), name "w_a_b"
inline bool xnode_less(const XNode &x, const XNode &y)   // std::string order of the names
{
    if (x.w != y.w) return x.w < y.w;                    // digits sort before 'w'
    if (x.a != y.a) return dec_cmp(x.a, y.a, true) < 0;  // "a_..." : the number is followed by '_'
    return dec_cmp(x.b, y.b, false) < 0;                 // same first number: the second one ends the name
}

}  // namespace phi_adapter
#endif
