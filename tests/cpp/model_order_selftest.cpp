// Checks the integer comparators of integration/phi_model.hpp against the std::string comparisons they stand in for
// (the reference orders the nodes of its expanded graph by iterating std::map<std::string, ...>, ILP_index.cpp:1312-1317),
// and the open-addressing tables against std::map.  Compiled and run by tests/test_model_block.py; no Gurobi, no GPU.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>
#include "phi_model_tables.hpp"

static int sgn(int x) { return (x > 0) - (x < 0); }

int main()
{
    using namespace phi_adapter;
    std::vector<uint32_t> nums;
    for (uint32_t v = 0; v < 1300; ++v) nums.push_back(v);
    const uint32_t extra[] = {9999, 10000, 10001, 99999, 100000, 123456, 1234567, 999999999u, 1000000000u, 4294967295u, 429496729u, 42949672u};
    for (size_t i = 0; i < sizeof(extra) / sizeof(extra[0]); ++i) nums.push_back(extra[i]);
    long checked = 0;
    for (size_t i = 0; i < nums.size(); ++i)
        for (size_t j = 0; j < nums.size(); ++j) {
            const std::string a = std::to_string(nums[i]), b = std::to_string(nums[j]);
            if (sgn(dec_cmp(nums[i], nums[j], false)) != sgn(a.compare(b))) { printf("plain %s %s\n", a.c_str(), b.c_str()); return 1; }
            if (sgn(dec_cmp(nums[i], nums[j], true)) != sgn((a + "_7").compare(b + "_7"))) { printf("underscore %s %s\n", a.c_str(), b.c_str()); return 1; }
            ++checked;
        }
    // node names: A(v, i) = "v_i", W(u, v) = "w_u_v"
    std::vector<XNode> nodes;
    const uint32_t vs[] = {0, 1, 2, 9, 10, 11, 19, 100, 101, 109, 110, 1000, 1099, 12, 120, 121};
    for (uint32_t w = 0; w < 2; ++w)
        for (size_t i = 0; i < 16; ++i)
            for (size_t j = 0; j < 16; ++j) { XNode n; n.w = w; n.a = vs[i]; n.b = vs[j]; nodes.push_back(n); }
    for (size_t i = 0; i < nodes.size(); ++i)
        for (size_t j = 0; j < nodes.size(); ++j) {
            const XNode &x = nodes[i], &y = nodes[j];
            const std::string sx = (x.w ? "w_" : "") + std::to_string(x.a) + "_" + std::to_string(x.b), sy = (y.w ? "w_" : "") + std::to_string(y.a) + "_" + std::to_string(y.b);
            if (xnode_less(x, y) != (sx < sy)) { printf("node %s %s\n", sx.c_str(), sy.c_str()); return 1; }
            ++checked;
        }
    // tables against std::map
    EdgeVarTable t; PairIndex p;
    std::map<std::vector<int32_t>, int64_t> ref; std::map<std::pair<uint32_t, uint32_t>, uint32_t> pref;
    uint64_t x = 88172645463325252ull; int64_t next = 0;
    for (int it = 0; it < 400000; ++it) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const int32_t u = (int32_t)(x % 3000), v = (int32_t)((x >> 20) % 50), j = (int32_t)((x >> 40) % 7);
        std::vector<int32_t> key; key.push_back(u); key.push_back(v); key.push_back(j);
        const int64_t got = t.find_or_reserve(u, v, j, next);
        if (ref.count(key)) { if (got != ref[key] || t.find(u, v, j) != got) { printf("table hit\n"); return 1; } }
        else { if (got != -1 || t.find(u, v, j) != next) { printf("table miss\n"); return 1; } ref[key] = next++; }
        bool is_new; const uint32_t id = p.get((uint32_t)u, (uint32_t)v, &is_new);
        const std::pair<uint32_t, uint32_t> pk((uint32_t)u, (uint32_t)v);
        if (pref.count(pk)) { if (is_new || id != pref[pk]) { printf("pair hit\n"); return 1; } }
        else { if (!is_new || id != pref.size()) { printf("pair miss\n"); return 1; } pref[pk] = id; }
        uint32_t f; if (!p.find((uint32_t)u, (uint32_t)v, f) || f != id || p.find(70000u, (uint32_t)v, f)) { printf("pair find\n"); return 1; }
    }
    if (t.find(1, 2, 999) != -1) { printf("absent key found\n"); return 1; }
    printf("MODEL_ORDER_OK %ld comparisons, %zu edge keys, %zu pairs\n", checked, ref.size(), pref.size());
    return 0;
}
