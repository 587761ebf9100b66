// phi_model.hpp — SURVEY.md §8(f) row 3: the k-mer constraint block of the Gurobi model construction
// (/root/reference/src/ILP_index.cpp:782-880, both the ILP and the QP branch) built STRAIGHT from the grouped result of
// libphi_gpu_index.so.
//
// The reference first materialises Anchor_hits[count_sp_r][num_walks] (one std::vector per (rank, walk), 24 bytes each even
// when empty, plus one heap vector per anchor) and then, for every edge (u, v) under every anchor, does three lookups in a
// std::map<std::string, GRBVar> keyed by "u_j_v_j" (find, operator[] to insert, operator[] to read).  Here
//   * the groups of a rank are walked in place: the (walk, group) pairs of the rank are put in (walk, group) order by a
//     counting pass over the member walks — that IS the reference's (j, k) iteration order, k being the position of the
//     group among the groups that contain walk j;
//   * the edge variables are found through an open-addressing table keyed by the integers (u, v, j); the string name is
//     built once per NEW variable (Gurobi needs it, and the unchanged code after the block still looks variables up in
//     `vars` by name, so the map receives exactly one insertion per variable instead of three lookups per occurrence);
//   * every addVar / addConstr / addQConstr call is the reference's, in the reference's order, with the reference's names —
//     tests compare the recorded model dump byte for byte (tests/test_model_block.py on the CPU, tests/test_gpu_dropin.py
//     on the GPU).
//
// Header-only C++11; compiled into the reference's ILP_index.cpp after gurobi_c++.h and phi_adapter.hpp.
#ifndef PHI_MODEL_HPP
#define PHI_MODEL_HPP

#include "phi_adapter.hpp"

#include <algorithm>
#include <map>
#include <string>
#include <vector>

namespace phi_adapter {

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

// Replaces ILP_index.cpp:782-880.  `vars`, `Zvars`, `count_kmer_matches` are the reference's locals (:774-780).
inline void add_kmer_constraints(GRBModel &model, const phi_index_result *res, int32_t num_walks, int32_t k_mer, bool is_ilp, bool is_mixed,
                                 std::map<std::string, GRBVar> &vars, std::vector<GRBVar> &Zvars, int32_t &count_kmer_matches)
{
    fprintf(stderr, "[M::%s::%.3f*%.2f] %s model started\n", "ILP_function", realtime() - mg_realtime0, cputime() / (realtime() - mg_realtime0),
            is_ilp ? "ILP" : "QP");                                                                  // :784 / :830
    EdgeVarTable table;
    std::vector<GRBVar> pool;                          // edge variables in creation order
    std::vector<uint32_t> walk_cnt(num_walks + 1), pair_group;      // per rank: groups of every walk, in (walk, group) order
    std::vector<uint64_t> group_voff(1, 0);
    const int32_t count_sp_r = res->count_sp_r;
    uint64_t voff = 0;
    for (int32_t i = 0; i < count_sp_r; ++i) {
        const uint32_t g0 = res->rank_off[i], g1 = res->rank_off[i + 1];
        GRBQuadExpr q_expr;                            // QP: one quadratic expression per rank (:834)
        GRBLinExpr z_expr;
        int32_t temp = 0;
        if (g1 > g0) {
            // (walk, group) order of the rank's anchors: counting pass over the member walks; groups stay in key order inside a walk
            std::fill(walk_cnt.begin(), walk_cnt.end(), 0u);
            group_voff.resize(g1 - g0 + 1);
            for (uint32_t g = g0; g < g1; ++g) {
                group_voff[g - g0] = voff; voff += res->group_len[g];
                for (uint32_t m = res->group_member_off[g]; m < res->group_member_off[g + 1]; ++m) ++walk_cnt[member_walk(res, m) + 1];
            }
            group_voff[g1 - g0] = voff;
            for (int32_t j = 0; j < num_walks; ++j) walk_cnt[j + 1] += walk_cnt[j];                 // first pair of every walk
            const uint32_t n_pairs = walk_cnt[num_walks];
            pair_group.resize(n_pairs);
            std::vector<uint32_t> cursor(walk_cnt.begin(), walk_cnt.end() - 1);
            for (uint32_t g = g0; g < g1; ++g)
                for (uint32_t m = res->group_member_off[g]; m < res->group_member_off[g + 1]; ++m) pair_group[cursor[member_walk(res, m)]++] = g;
            for (int32_t j = 0; j < num_walks; ++j) {
                for (uint32_t p = walk_cnt[j]; p < walk_cnt[j + 1]; ++p) {
                    const int32_t k = (int32_t)(p - walk_cnt[j]);
                    const uint32_t g = pair_group[p];
                    const int32_t *list = res->group_vtx + group_voff[g - g0];
                    const int32_t n = res->group_len[g];
                    GRBLinExpr kmer_expr;              // ILP: one linear expression per anchor (:792)
                    const std::string extra_var = "z_" + std::to_string(i) + "_" + std::to_string(j) + "_" + std::to_string(k);
                    GRBVar kmer_expr_var = model.addVar(0.0, 1.0, 0.0, GRB_BINARY, extra_var);
                    if (n - 1 == 0) continue;          // ignore matches with only one vertex (:795 / :841)
                    if (!is_ilp) { const int32_t weight = (k_mer - 1) - (n - 1); q_expr += weight * kmer_expr_var; }   // :842-843
                    for (int32_t l = 1; l < n; ++l) {
                        const int32_t u = list[l - 1], v = list[l];
                        int64_t idx = table.find_or_reserve(u, v, j, (int64_t)pool.size());
                        if (idx < 0) {                 // variable does not exist (:802-812 / :848-857)
                            const std::string var_name = std::to_string(u) + "_" + std::to_string(j) + "_" + std::to_string(v) + "_" + std::to_string(j);
                            idx = (int64_t)pool.size();
                            pool.push_back(model.addVar(0.0, 1.0, 0.0, is_mixed ? GRB_CONTINUOUS : GRB_BINARY, var_name));
                            vars[var_name] = pool.back();
                        }
                        if (is_ilp) kmer_expr += pool[idx];                                          // :814
                        else q_expr += pool[idx] * kmer_expr_var;                                    // :858
                    }
                    if (is_ilp) {
                        const int32_t weight = n - 1;                                                // :816-817
                        model.addConstr(kmer_expr >= weight * kmer_expr_var,
                                        "Kmer_constraints_" + std::to_string(i) + "_" + std::to_string(j) + "_" + std::to_string(k));
                    }
                    z_expr += kmer_expr_var;
                    temp += 1;
                }
            }
        }
        if (temp != 0) {                               // :822-832 / :864-874
            const std::string constraint_name = "Kmer_constraints_" + std::to_string(i);
            const int32_t kmer_weight = k_mer - 1;
            const std::string z_var = "z_" + std::to_string(i);
            GRBVar z_var_r = model.addVar(0.0, 1.0, 0.0, GRB_BINARY, z_var);
            Zvars.push_back(z_var_r);
            if (!is_ilp) model.addQConstr(q_expr == kmer_weight * z_var_r, constraint_name);
            model.addConstr(z_expr == z_var_r, "Z_constraint_" + std::to_string(i));
            count_kmer_matches++;
        }
    }
}

}  // namespace phi_adapter
#endif
