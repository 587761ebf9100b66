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
#include "phi_model_tables.hpp"

#include <algorithm>
#include <map>
#include <string>
#include <vector>

namespace phi_adapter {

// What the blocks share: every edge variable u_j_v_j created so far (the string map `vars` of the reference stays in step for the
// code that is not replaced).
struct ModelState {
    EdgeVarTable same_walk;            // (u, v, j)  ->  variable "u_j_v_j"
    std::vector<GRBVar> pool;          // all variables created through the tables, in creation order
};

// Replaces ILP_index.cpp:782-880.  `vars`, `Zvars`, `count_kmer_matches` are the reference's locals (:774-780).
inline void add_kmer_constraints(GRBModel &model, ModelState &st, const FrontEnd &fe, int32_t num_walks, int32_t k_mer, bool is_ilp, bool is_mixed,
                                 std::map<std::string, GRBVar> &vars, std::vector<GRBVar> &Zvars, int32_t &count_kmer_matches)
{
    fprintf(stderr, "[M::%s::%.3f*%.2f] %s model started\n", "ILP_function", realtime() - mg_realtime0, cputime() / (realtime() - mg_realtime0),
            is_ilp ? "ILP" : "QP");                                                                  // :784 / :830
    EdgeVarTable &table = st.same_walk;
    std::vector<GRBVar> &pool = st.pool;               // edge variables in creation order
    std::vector<uint32_t> walk_cnt(num_walks + 1), pair_group;      // per rank: groups of every walk, in (walk, group) order
    std::vector<uint64_t> group_voff(1, 0);
    const int32_t count_sp_r = fe.count_sp_r();
    std::vector<uint64_t> part_voff(fe.parts.size(), 0);            // running vertex offset inside every part
    for (int32_t i = 0; i < count_sp_r; ++i) {
        GRBQuadExpr q_expr;                            // QP: one quadratic expression per rank (:834)
        GRBLinExpr z_expr;
        int32_t temp = 0;
        // several GPUs: part p holds the groups of its own walks, and walk ranges ascend with p — the parts one after another
        // give the reference's ascending j; the k of an anchor only counts the groups of its own walk, which live in one part
        for (size_t part = 0; part < fe.parts.size(); ++part) {
        const phi_index_result *res = fe.parts[part];
        uint64_t &voff = part_voff[part];
        const uint32_t g0 = res->rank_off[i], g1 = res->rank_off[i + 1];
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
                    const std::string extra_var = Name().s("z_").i(i).c('_').i(j).c('_').i(k).str();
                    GRBVar kmer_expr_var = model.addVar(0.0, 1.0, 0.0, GRB_BINARY, extra_var);
                    if (n - 1 == 0) continue;          // ignore matches with only one vertex (:795 / :841)
                    if (!is_ilp) { const int32_t weight = (k_mer - 1) - (n - 1); q_expr += weight * kmer_expr_var; }   // :842-843
                    for (int32_t l = 1; l < n; ++l) {
                        const int32_t u = list[l - 1], v = list[l];
                        int64_t idx = table.find_or_reserve(u, v, j, (int64_t)pool.size());
                        if (idx < 0) {                 // variable does not exist (:802-812 / :848-857)
                            const std::string var_name = Name().i(u).c('_').i(j).c('_').i(v).c('_').i(j).str();
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
                                        Name().s("Kmer_constraints_").i(i).c('_').i(j).c('_').i(k).str());
                    }
                    z_expr += kmer_expr_var;
                    temp += 1;
                }
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

// ------------------------------------------------------------------------------------------------------------------
// The "optimized expanded graph" part of the model (the default branch, /root/reference/src/ILP_index.cpp:1201-1406): edge
// variables without recombination, the recombination vertices w_u_v with their edges, the objective, and the flow
// conservation constraints.  The reference keeps the expanded graph in std::map<std::string, std::vector<std::string>>
// (new_adj, in_nodes_new) and finds every variable by a freshly concatenated name in std::map<std::string, GRBVar>; that is
// where its 6-7 s on the README data go.  Here nodes are integers (A(v, i) = vertex v on walk i, W(u, v) = the recombination
// vertex of the edge u -> v), adjacency lists are vectors, variables are found through integer-keyed tables, and the ONE place
// where the reference's result depends on string order — in_nodes_new is filled by iterating new_adj in key order, so the
// predecessors of a node come in the std::string order of their names — is reproduced by sorting every predecessor list with
// an integer comparator that orders (v, i) / (u, v) exactly as "v_i" / "w_u_v" compare as strings.
// Every model call is the reference's, in its order, with its names (tests/test_model_block.py, tests/test_gpu_dropin.py).

// Replaces ILP_index.cpp:1201-1406.  vtx_expr / obj are the reference's locals (:1163, :1197); c_1 = recombination (:776).
inline void add_expanded_graph(GRBModel &model, ModelState &st, ILP_index &ix, bool is_mixed, int32_t c_1,
                               std::map<std::string, GRBVar> &vars, std::vector<GRBVar> &Zvars, GRBLinExpr &vtx_expr, GRBLinExpr &obj)
{
    const char vtype = is_mixed ? GRB_CONTINUOUS : GRB_BINARY;
    const int32_t num_walks = (int32_t)ix.num_walks;
    std::vector<GRBVar> &pool = st.pool;
    EdgeVarTable to_w, from_w;                           // (u, v, h) -> "u_h_w_u_v",  (u, v, h) -> "w_u_v_v_h"
    PairIndex a_index, w_index;                          // A(v, i), W(u, v) -> dense ids
    std::vector<XNode> a_node, w_node;
    std::vector<std::vector<uint32_t> > a_adj, w_adj;    // successors; bit 31 set: a W node
    std::vector<char> w_is_key;                          // W(u, v) has an entry in new_adj (it has successors)
    const uint32_t WBIT = 0x80000000u;
    struct Ids {
        static uint32_t a(PairIndex &ai, std::vector<XNode> &nodes, std::vector<std::vector<uint32_t> > &adj, uint32_t v, uint32_t i)
        {
            bool is_new; const uint32_t id = ai.get(v, i, &is_new);
            if (is_new) { XNode n; n.w = 0; n.a = v; n.b = i; nodes.push_back(n); adj.push_back(std::vector<uint32_t>()); }
            return id;
        }
    };

    // w/o recombination (:1202-1226)
    for (int32_t i = 0; i < num_walks; i++) {
        for (size_t idx = 0; idx + 1 < ix.paths[i].size(); idx++) {
            const int32_t u = ix.paths[i][idx], v = ix.paths[i][idx + 1];
            const uint32_t au = Ids::a(a_index, a_node, a_adj, u, i), av = Ids::a(a_index, a_node, a_adj, v, i);
            a_adj[au].push_back(av);
            if (st.same_walk.find_or_reserve(u, v, i, (int64_t)pool.size()) < 0) {      // variable does not exist
                const std::string var_name = Name().i(u).c('_').i(i).c('_').i(v).c('_').i(i).str();
                pool.push_back(model.addVar(0.0, 1.0, 0.0, vtype, var_name));
                vtx_expr += 0 * pool.back();             // no need without recombination
            }
        }
    }
    if (getenv("PHI_MODEL_TIMES")) fprintf(stderr, "[phi_model] %.3f same-walk edges\n", realtime() - mg_realtime0);
    // index of a vertex in a haplotype, last occurrence (:1229-1238)
    PairIndex pos_index; std::vector<uint32_t> pos_of;
    for (size_t h = 0; h < ix.paths.size(); h++)
        for (size_t i = 0; i < ix.paths[h].size(); ++i) {
            const uint32_t id = pos_index.get(ix.paths[h][i], (uint32_t)h);
            if (id == pos_of.size()) pos_of.push_back((uint32_t)i); else pos_of[id] = (uint32_t)i;
        }
    // recombination vertices and edges (:1241-1298)
    for (uint32_t u = 0; u < ix.adj_list.size(); u++) {
        for (size_t q = 0; q < ix.adj_list[u].size(); ++q) {
            const uint32_t v = ix.adj_list[u][q];
            bool new_vertex_used = false;
            uint32_t wid = 0;
            std::string new_vtx;
            for (size_t t = 0; t < ix.haps[u].size(); ++t) {
                const uint32_t h = ix.haps[u][t];
                uint32_t pid; int index = 0;
                if (pos_index.find(u, h, pid)) index = (int)pos_of[pid];                 // (a missing entry reads as 0 in the reference's map)
                if (index == (int)ix.paths[h].size() - 1 || ix.paths[h][index + 1] != v) {
                    if (!new_vertex_used) {
                        new_vertex_used = true;
                        bool is_new; wid = w_index.get(u, v, &is_new);
                        if (is_new) { XNode n; n.w = 1; n.a = u; n.b = v; w_node.push_back(n); w_adj.push_back(std::vector<uint32_t>()); w_is_key.push_back(0); }
                        new_vtx = Name().s("w_").i(u).c('_').i(v).str();
                    }
                    a_adj[Ids::a(a_index, a_node, a_adj, u, h)].push_back(wid | WBIT);
                    int64_t idx = to_w.find_or_reserve((int32_t)u, (int32_t)v, (int32_t)h, (int64_t)pool.size());
                    if (idx < 0) {
                        idx = (int64_t)pool.size();
                        pool.push_back(model.addVar(0.0, 1.0, 0.0, vtype, Name().i(u).c('_').i(h).c('_').s(new_vtx.c_str()).str()));
                    }
                    vtx_expr += (c_1 / 2) * pool[idx];
                }
            }
            if (new_vertex_used) {
                for (size_t t = 0; t < ix.haps[v].size(); ++t) {
                    const uint32_t h = ix.haps[v][t];
                    w_adj[wid].push_back(Ids::a(a_index, a_node, a_adj, v, h)); w_is_key[wid] = 1;
                    int64_t idx = from_w.find_or_reserve((int32_t)u, (int32_t)v, (int32_t)h, (int64_t)pool.size());
                    if (idx < 0) {
                        idx = (int64_t)pool.size();
                        pool.push_back(model.addVar(0.0, 1.0, 0.0, vtype, Name().s(new_vtx.c_str()).c('_').i(v).c('_').i(h).str()));
                    }
                    vtx_expr += (c_1 / 2) * pool[idx];
                }
            }
        }
    }
    if (getenv("PHI_MODEL_TIMES")) fprintf(stderr, "[phi_model] %.3f recombination edges\n", realtime() - mg_realtime0);
    // (1 - z_i) terms, objective (:1301-1309)
    GRBLinExpr z_expr;
    for (size_t i = 0; i < Zvars.size(); i++) z_expr += (1 - Zvars[i]);
    obj = vtx_expr + z_expr;
    model.setObjective(obj, GRB_MINIMIZE);

    if (getenv("PHI_MODEL_TIMES")) fprintf(stderr, "[phi_model] %.3f objective\n", realtime() - mg_realtime0);
    // reverse adjacency (:1312-1317): the reference iterates new_adj in key order, so every predecessor list is in the string
    // order of the predecessors' names (equal names stay together)
    std::vector<std::vector<uint32_t> > a_in(a_node.size()), w_in(w_node.size());
    for (uint32_t s = 0; s < a_adj.size(); ++s)
        for (size_t q = 0; q < a_adj[s].size(); ++q) { const uint32_t t = a_adj[s][q]; if (t & WBIT) w_in[t & ~WBIT].push_back(s); else a_in[t].push_back(s); }
    for (uint32_t s = 0; s < w_adj.size(); ++s)
        for (size_t q = 0; q < w_adj[s].size(); ++q) a_in[w_adj[s][q]].push_back(s | WBIT);
    struct ByName {
        const std::vector<XNode> *an, *wn;
        bool operator()(uint32_t x, uint32_t y) const
        {
            const XNode &nx = (x & 0x80000000u) ? (*wn)[x & 0x7FFFFFFFu] : (*an)[x], &ny = (y & 0x80000000u) ? (*wn)[y & 0x7FFFFFFFu] : (*an)[y];
            return xnode_less(nx, ny);
        }
    } by_name; by_name.an = &a_node; by_name.wn = &w_node;
    for (size_t t = 0; t < a_in.size(); ++t) if (a_in[t].size() > 1) std::stable_sort(a_in[t].begin(), a_in[t].end(), by_name);
    for (size_t t = 0; t < w_in.size(); ++t) if (w_in[t].size() > 1) std::stable_sort(w_in[t].begin(), w_in[t].end(), by_name);

    // variable of the edge s -> t of the expanded graph
    struct EdgeVar {
        ModelState *st; EdgeVarTable *to_w, *from_w; const std::vector<XNode> *an, *wn;
        GRBVar &operator()(uint32_t s, uint32_t t) const
        {
            int64_t idx;
            if (t & 0x80000000u) { const XNode &a = (*an)[s], &w = (*wn)[t & 0x7FFFFFFFu]; idx = to_w->find((int32_t)w.a, (int32_t)w.b, (int32_t)a.b); }
            else if (s & 0x80000000u) { const XNode &w = (*wn)[s & 0x7FFFFFFFu], &a = (*an)[t]; idx = from_w->find((int32_t)w.a, (int32_t)w.b, (int32_t)a.b); }
            else { const XNode &a = (*an)[s], &b = (*an)[t]; idx = st->same_walk.find((int32_t)a.a, (int32_t)b.a, (int32_t)a.b); }
            if (idx < 0) { fprintf(stderr, "Error: expanded-graph edge without a variable\n"); exit(1); }
            return st->pool[idx];
        }
    } edge_var; edge_var.st = &st; edge_var.to_w = &to_w; edge_var.from_w = &from_w; edge_var.an = &a_node; edge_var.wn = &w_node;

    if (getenv("PHI_MODEL_TIMES")) fprintf(stderr, "[phi_model] %.3f reverse adjacency\n", realtime() - mg_realtime0);
    // paths based flow constraints (:1320-1343)
    for (int32_t i = 0; i < num_walks; i++) {
        for (size_t idx = 0; idx < ix.paths[i].size(); idx++) {
            if (idx == 0 || idx == ix.paths[i].size() - 1) continue;                     // skip source and sink nodes
            GRBLinExpr in_expr, out_expr;
            const int32_t v = ix.paths[i][idx];
            const uint32_t t = Ids::a(a_index, a_node, a_adj, v, i);
            if (t >= a_in.size()) a_in.resize(a_node.size());
            for (size_t q = 0; q < a_in[t].size(); ++q) in_expr += edge_var(a_in[t][q], t);
            for (size_t q = 0; q < a_adj[t].size(); ++q) out_expr += edge_var(t, a_adj[t][q]);
            model.addConstr(in_expr == out_expr, Name().s("Flow_conservation_").i(v).c('_').i(i).str());
        }
    }
    if (getenv("PHI_MODEL_TIMES")) fprintf(stderr, "[phi_model] %.3f path flow constraints\n", realtime() - mg_realtime0);
    // w_u_v vertices (:1345-1368)
    for (uint32_t u = 0; u < ix.n_vtx; u++) {
        for (size_t q = 0; q < ix.adj_list[u].size(); ++q) {
            const uint32_t v = ix.adj_list[u][q];
            uint32_t wid;
            if (!w_index.find(u, v, wid) || !w_is_key[wid]) continue;                    // w_vtx exists
            GRBLinExpr in_expr, out_expr;
            for (size_t r = 0; r < w_adj[wid].size(); ++r) out_expr += edge_var(wid | WBIT, w_adj[wid][r]);
            for (size_t r = 0; r < w_in[wid].size(); ++r) in_expr += edge_var(w_in[wid][r], wid | WBIT);
            model.addConstr(in_expr == out_expr, Name().s("Flow_conservation_w_").i(u).c('_').i(v).str());
        }
    }
    if (getenv("PHI_MODEL_TIMES")) fprintf(stderr, "[phi_model] %.3f w flow constraints\n", realtime() - mg_realtime0);
    // source nodes (:1371-1382)
    for (int32_t i = 0; i < num_walks; i++) {
        const int32_t u = ix.paths[i][0];
        GRBLinExpr s_expr;
        s_expr += vars["s_" + std::to_string(u) + "_" + std::to_string(i)];
        const uint32_t t = Ids::a(a_index, a_node, a_adj, u, i);
        for (size_t q = 0; q < a_adj[t].size(); ++q) s_expr -= edge_var(t, a_adj[t][q]);
        model.addConstr(s_expr == 0, "Source_conservation_" + std::to_string(u) + "_" + std::to_string(i));
    }
    // sink nodes (:1385-1398)
    for (int32_t i = 0; i < num_walks; i++) {
        const int32_t u = ix.paths[i].back();
        GRBLinExpr e_expr;
        const uint32_t t = Ids::a(a_index, a_node, a_adj, u, i);
        if (t >= a_in.size()) a_in.resize(a_node.size());
        for (size_t q = 0; q < a_in[t].size(); ++q) e_expr += edge_var(a_in[t][q], t);
        e_expr += -1 * vars[std::to_string(u) + "_" + std::to_string(i) + "_e"];
        model.addConstr(e_expr == 0, "Sink_conservation_" + std::to_string(u) + "_" + std::to_string(i));
    }
    vars.clear();                                                                        // :1402
}

// ------------------------------------------------------------------------------------------------------------------
// The "naive expanded graph" part of the model (-N1, /root/reference/src/ILP_index.cpp:942-1154): every edge variable
// u_i_v_j (vertex u on walk i -> vertex v on walk j), the objective and the flow conservation constraints.  The reference's
// only string-keyed structure here is `vars`; it is replaced by two integer-keyed tables (same walk: the one the k-mer block
// filled, (u, v, i); two walks: (u, i, v, j)).  Loops, call order and names are the reference's.
inline void add_naive_graph(GRBModel &model, ModelState &st, ILP_index &ix, bool is_mixed, int32_t c_1,
                            const std::vector<std::vector<int32_t> > &in_nodes, std::map<std::string, GRBVar> &vars, std::vector<GRBVar> &Zvars,
                            GRBLinExpr &vtx_expr, GRBLinExpr &obj)
{
    const char vtype = is_mixed ? GRB_CONTINUOUS : GRB_BINARY;
    const int32_t num_walks = (int32_t)ix.num_walks;
    std::vector<GRBVar> &pool = st.pool;
    CrossVarTable cross;
    struct Vars {
        GRBModel *model; ModelState *st; CrossVarTable *cross; char vtype;
        // the variable "u_i_v_j"; created (and *created set) when it does not exist yet
        GRBVar &get(int32_t u, int32_t i, int32_t v, int32_t j, bool *created = 0)
        {
            std::vector<GRBVar> &pool = st->pool;
            int64_t idx = i == j ? st->same_walk.find_or_reserve(u, v, i, (int64_t)pool.size()) : cross->find_or_reserve(u, i, v, j, (int64_t)pool.size());
            if (created) *created = idx < 0;
            if (idx < 0) {
                idx = (int64_t)pool.size();
                pool.push_back(model->addVar(0.0, 1.0, 0.0, vtype, Name().i(u).c('_').i(i).c('_').i(v).c('_').i(j).str()));
            }
            return pool[idx];
        }
    } V; V.model = &model; V.st = &st; V.cross = &cross; V.vtype = vtype;
    (void)pool;

    // w/o recombination (:942-962)
    for (int32_t i = 0; i < num_walks; i++)
        for (size_t idx = 0; idx + 1 < ix.paths[i].size(); idx++) {
            bool created; GRBVar &var = V.get(ix.paths[i][idx], i, ix.paths[i][idx + 1], i, &created);
            if (created) vtx_expr += 0 * var;                                              // no need without recombination
        }
    // with recombination (:965-995)
    for (int32_t i = 0; i < num_walks; i++)
        for (size_t idx = 0; idx + 1 < ix.paths[i].size(); idx++) {
            const int32_t u = ix.paths[i][idx];
            for (size_t a = 0; a < ix.adj_list[u].size(); ++a) {
                const int32_t v = ix.adj_list[u][a];
                for (size_t b = 0; b < ix.haps[v].size(); ++b) {
                    const int32_t j = ix.haps[v][b];
                    if (i == j) continue;
                    bool created; GRBVar &var = V.get(u, i, v, j, &created);
                    if (created) vtx_expr += c_1 * var;
                }
            }
        }
    // (1 - z_i) terms, objective (:998-1004)
    GRBLinExpr z_expr;
    for (size_t i = 0; i < Zvars.size(); i++) z_expr += (1 - Zvars[i]);
    obj = vtx_expr + z_expr;
    model.setObjective(obj, GRB_MINIMIZE);
    // paths based flow constraints (:1007-1088)
    for (int32_t i = 0; i < num_walks; i++)
        for (size_t idx = 0; idx < ix.paths[i].size(); idx++) {
            if (idx == 0 || idx == ix.paths[i].size() - 1) continue;                       // skip source and sink nodes
            GRBLinExpr in_expr, out_expr;
            const int32_t v = ix.paths[i][idx], v_in = ix.paths[i][idx - 1], v_out = ix.paths[i][idx + 1];
            in_expr += V.get(v_in, i, v, i);
            out_expr += V.get(v, i, v_out, i);
            for (size_t a = 0; a < in_nodes[v].size(); ++a) {                              // in expression
                const int32_t u = in_nodes[v][a];
                for (size_t b = 0; b < ix.haps[u].size(); ++b) { const int32_t j = ix.haps[u][b]; if (i != j) in_expr += V.get(u, j, v, i); }
            }
            for (size_t a = 0; a < ix.adj_list[v].size(); ++a) {                           // out expression
                const int32_t u = ix.adj_list[v][a];
                for (size_t b = 0; b < ix.haps[u].size(); ++b) { const int32_t j = ix.haps[u][b]; if (i != j) out_expr += V.get(v, i, u, j); }
            }
            model.addConstr(in_expr == out_expr, Name().s("Flow_conservation_").i(v).c('_').i(i).str());
        }
    // source nodes (:1092-1113)
    for (int32_t i = 0; i < num_walks; i++) {
        const int32_t u = ix.paths[i][0];
        GRBLinExpr s_expr;
        s_expr += vars["s_" + std::to_string(u) + "_" + std::to_string(i)];
        for (size_t a = 0; a < ix.adj_list[u].size(); ++a) {
            const int32_t v = ix.adj_list[u][a];
            for (size_t b = 0; b < ix.haps[v].size(); ++b) s_expr -= V.get(u, i, v, (int32_t)ix.haps[v][b]);
        }
        model.addConstr(s_expr == 0, "Source_conservation_" + std::to_string(u) + "_" + std::to_string(i));
    }
    // sink nodes (:1116-1153)
    for (int32_t i = 0; i < num_walks; i++) {
        const int32_t u = ix.paths[i][ix.paths[i].size() - 1];
        GRBLinExpr e_expr;
        for (size_t a = 0; a < in_nodes[u].size(); ++a) {
            const int32_t v = in_nodes[u][a];
            for (size_t b = 0; b < ix.haps[v].size(); ++b) e_expr += V.get(v, (int32_t)ix.haps[v][b], u, i);
        }
        const std::string var_name = std::to_string(u) + "_" + std::to_string(i) + "_e";
        if (vars.find(var_name) == vars.end()) vars[var_name] = model.addVar(0.0, 1.0, 0.0, vtype, var_name);   // (it exists: created with the start variables)
        e_expr += -1 * vars[var_name];
        model.addConstr(e_expr == 0, "Sink_conservation_" + std::to_string(u) + "_" + std::to_string(i));
    }
}

}  // namespace phi_adapter
#endif
