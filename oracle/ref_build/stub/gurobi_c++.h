// TEST INFRASTRUCTURE ONLY — recording stand-in for Gurobi's C++ header.
//
// Gurobi 11.0.2 is proprietary and absent from this image, but the reference
// (/root/reference/src/ILP_index.h:38) includes "gurobi_c++.h".  This header
// provides just the API subset that /root/reference/src/ILP_index.cpp uses
// (lines 157-310 printers, 757-1442 model construction / solve / back-trace)
// so the UNMODIFIED reference sources compile.  Instead of building a solver
// model it records every addVar / addConstr / addQConstr / setObjective call,
// in creation order, into a text dump (path: $PHI_STUB_DUMP, default
// "model_dump.txt").  optimize() flushes the dump and throws GRBException,
// which the reference catches (ILP_index.cpp:1583) before writing an empty
// FASTA.  Nothing in the product (phi_b200/) may include this file.
//
// Dump grammar (one record per line, creation order preserved):
//   P <name> <value>                              model.set(string,string) / env.set
//   V <name> <B|C> <lb> <ub> <obj>                addVar
//   C <name> <sense> | <lhs terms> | <rhs terms>  addConstr   (terms: coeff*var, const as coeff*1)
//   Q <name> <sense> | <lhs lin> ; <lhs quad> | <rhs lin> ; <rhs quad>
//   O <sense> | <lin terms>                       setObjective
// Terms are NOT merged or reordered: they appear exactly as the reference
// appended them, so the dump pins ordering as well as content.
#ifndef PHI_ORACLE_GUROBI_STUB_H
#define PHI_ORACLE_GUROBI_STUB_H

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <utility>

#define GRB_BINARY 'B'
#define GRB_CONTINUOUS 'C'
#define GRB_INTEGER 'I'
#define GRB_EQUAL '='
#define GRB_LESS_EQUAL '<'
#define GRB_GREATER_EQUAL '>'
#define GRB_MINIMIZE 1
#define GRB_MAXIMIZE -1

enum GRB_IntParam { GRB_IntParam_Threads };
enum GRB_DoubleParam { GRB_DoubleParam_OptimalityTol };
enum GRB_IntAttr { GRB_IntAttr_NumQConstrs, GRB_IntAttr_NumConstrs, GRB_IntAttr_NumVars, GRB_IntAttr_ModelSense };
enum GRB_StringAttr { GRB_StringAttr_QCName, GRB_StringAttr_VarName, GRB_StringAttr_ConstrName };
enum GRB_DoubleAttr { GRB_DoubleAttr_QCRHS, GRB_DoubleAttr_RHS, GRB_DoubleAttr_X, GRB_DoubleAttr_ObjCon };
enum GRB_CharAttr { GRB_CharAttr_QCSense, GRB_CharAttr_Sense, GRB_CharAttr_VType };

class GRBException {
    std::string msg_; int code_;
public:
    GRBException(const std::string &m = "stub", int c = 0) : msg_(m), code_(c) {}
    int getErrorCode() const { return code_; }
    std::string getMessage() const { return msg_; }
};

struct phi_stub_state {
    std::vector<std::string> var_names;
    FILE *fp;
    long n_vars, n_lin, n_quad;
    bool quiet;                       // PHI_STUB_QUIET=1: count the calls, write nothing (timing of the caller's own model construction)
    phi_stub_state() : fp(0), n_vars(0), n_lin(0), n_quad(0), quiet(getenv("PHI_STUB_QUIET") != 0) {}
    FILE *out() {
        if (!fp) {
            const char *p = getenv("PHI_STUB_DUMP");
            fp = fopen(p ? p : "model_dump.txt", "w");
            if (!fp) { perror("PHI_STUB_DUMP"); exit(2); }
        }
        return fp;
    }
};
inline phi_stub_state &phi_stub() { static phi_stub_state s; return s; }

class GRBVar {
public:
    int id;
    GRBVar() : id(-1) {}
    explicit GRBVar(int i) : id(i) {}
    double get(GRB_DoubleAttr) const { return 0.0; }
    std::string get(GRB_StringAttr) const { return id >= 0 ? phi_stub().var_names[id] : std::string(); }
    char get(GRB_CharAttr) const { return 'C'; }
};

class GRBLinExpr {
public:
    std::vector<std::pair<double, int> > t;  // (coeff, var id); var id -1 == constant
    GRBLinExpr() {}
    GRBLinExpr(double c) { if (c != 0.0) t.push_back(std::make_pair(c, -1)); }
    GRBLinExpr(const GRBVar &v) { t.push_back(std::make_pair(1.0, v.id)); }
    GRBLinExpr(const GRBVar &v, double c) { t.push_back(std::make_pair(c, v.id)); }
    unsigned int size() const { unsigned n = 0; for (size_t i = 0; i < t.size(); ++i) n += t[i].second >= 0; return n; }
    GRBVar getVar(int i) const { return GRBVar(t[i].second); }
    double getCoeff(int i) const { return t[i].first; }
    GRBLinExpr &operator+=(const GRBLinExpr &o) { t.insert(t.end(), o.t.begin(), o.t.end()); return *this; }
    GRBLinExpr &operator-=(const GRBLinExpr &o) {
        for (size_t i = 0; i < o.t.size(); ++i) t.push_back(std::make_pair(-o.t[i].first, o.t[i].second));
        return *this;
    }
};
inline GRBLinExpr operator+(GRBLinExpr a, const GRBLinExpr &b) { a += b; return a; }
inline GRBLinExpr operator-(GRBLinExpr a, const GRBLinExpr &b) { a -= b; return a; }
inline GRBLinExpr operator*(double c, const GRBVar &v) { return GRBLinExpr(v, c); }
inline GRBLinExpr operator*(const GRBVar &v, double c) { return GRBLinExpr(v, c); }
inline GRBLinExpr operator-(double c, const GRBVar &v) { GRBLinExpr e(c); e -= GRBLinExpr(v); return e; }
inline GRBLinExpr operator+(double c, const GRBVar &v) { GRBLinExpr e(c); e += GRBLinExpr(v); return e; }

class GRBQuadExpr {
public:
    GRBLinExpr lin;
    struct qterm { double c; int a, b; };
    std::vector<qterm> q;
    GRBQuadExpr() {}
    GRBQuadExpr(const GRBLinExpr &l) : lin(l) {}
    GRBQuadExpr(double c) : lin(c) {}
    GRBQuadExpr(const GRBVar &v) : lin(v) {}
    unsigned int size() const { return (unsigned)q.size(); }
    GRBLinExpr getLinExpr() const { return lin; }
    GRBVar getVar1(int i) const { return GRBVar(q[i].a); }
    GRBVar getVar2(int i) const { return GRBVar(q[i].b); }
    double getCoeff(int i) const { return q[i].c; }
    GRBQuadExpr &operator+=(const GRBQuadExpr &o) { lin += o.lin; q.insert(q.end(), o.q.begin(), o.q.end()); return *this; }
};
inline GRBQuadExpr operator*(const GRBVar &a, const GRBVar &b) {
    GRBQuadExpr e; GRBQuadExpr::qterm t; t.c = 1.0; t.a = a.id; t.b = b.id; e.q.push_back(t); return e;
}

class GRBTempConstr {
public:
    GRBQuadExpr lhs, rhs; char sense;
    GRBTempConstr(const GRBQuadExpr &l, char s, const GRBQuadExpr &r) : lhs(l), rhs(r), sense(s) {}
};
inline GRBTempConstr operator==(const GRBQuadExpr &l, const GRBQuadExpr &r) { return GRBTempConstr(l, GRB_EQUAL, r); }
inline GRBTempConstr operator>=(const GRBQuadExpr &l, const GRBQuadExpr &r) { return GRBTempConstr(l, GRB_GREATER_EQUAL, r); }
inline GRBTempConstr operator<=(const GRBQuadExpr &l, const GRBQuadExpr &r) { return GRBTempConstr(l, GRB_LESS_EQUAL, r); }

class GRBConstr {
public:
    std::string get(GRB_StringAttr) const { return ""; }
    double get(GRB_DoubleAttr) const { return 0.0; }
    char get(GRB_CharAttr) const { return '='; }
};
class GRBQConstr {
public:
    std::string get(GRB_StringAttr) const { return ""; }
    double get(GRB_DoubleAttr) const { return 0.0; }
    char get(GRB_CharAttr) const { return '='; }
};

class GRBEnv {
public:
    GRBEnv(bool = false) {}
    void set(GRB_IntParam, int) {}   // thread count is not part of the model; not dumped
    void start() {}
};

class GRBModel {
    static void put_lin(FILE *f, const GRBLinExpr &e) {
        for (size_t i = 0; i < e.t.size(); ++i) {
            if (e.t[i].second < 0) fprintf(f, " %.17g*1", e.t[i].first);
            else fprintf(f, " %.17g*%s", e.t[i].first, phi_stub().var_names[e.t[i].second].c_str());
        }
    }
    static void put_quad(FILE *f, const GRBQuadExpr &e) {
        for (size_t i = 0; i < e.q.size(); ++i)
            fprintf(f, " %.17g*%s*%s", e.q[i].c, phi_stub().var_names[e.q[i].a].c_str(), phi_stub().var_names[e.q[i].b].c_str());
    }
public:
    GRBModel(const GRBEnv &) {}
    void set(const std::string &k, const std::string &v) { fprintf(phi_stub().out(), "P %s %s\n", k.c_str(), v.c_str()); }
    void set(GRB_DoubleParam, double v) { fprintf(phi_stub().out(), "P OptimalityTol %.17g\n", v); }
    double get(GRB_DoubleParam) const { return 1e-6; }
    int get(GRB_IntAttr a) const {
        if (a == GRB_IntAttr_NumVars) return (int)phi_stub().n_vars;
        if (a == GRB_IntAttr_NumConstrs) return (int)phi_stub().n_lin;
        if (a == GRB_IntAttr_NumQConstrs) return (int)phi_stub().n_quad;
        return GRB_MINIMIZE;
    }
    double get(GRB_DoubleAttr) const { return 0.0; }
    GRBVar addVar(double lb, double ub, double obj, char type, const std::string &name) {
        phi_stub_state &s = phi_stub();
        s.var_names.push_back(name);
        s.n_vars++;
        if (!s.quiet) fprintf(s.out(), "V %s %c %.17g %.17g %.17g\n", name.c_str(), type, lb, ub, obj);
        return GRBVar((int)s.var_names.size() - 1);
    }
    GRBConstr addConstr(const GRBTempConstr &c, const std::string &name) {
        phi_stub().n_lin++;
        if (phi_stub().quiet) return GRBConstr();
        FILE *f = phi_stub().out();
        fprintf(f, "C %s %c |", name.c_str(), c.sense);
        put_lin(f, c.lhs.lin); fputs(" |", f); put_lin(f, c.rhs.lin); fputc('\n', f);
        return GRBConstr();
    }
    GRBQConstr addQConstr(const GRBTempConstr &c, const std::string &name) {
        phi_stub().n_quad++;
        if (phi_stub().quiet) return GRBQConstr();
        FILE *f = phi_stub().out();
        fprintf(f, "Q %s %c |", name.c_str(), c.sense);
        put_lin(f, c.lhs.lin); fputs(" ;", f); put_quad(f, c.lhs);
        fputs(" |", f); put_lin(f, c.rhs.lin); fputs(" ;", f); put_quad(f, c.rhs); fputc('\n', f);
        return GRBQConstr();
    }
    void setObjective(const GRBLinExpr &e, int sense) {
        if (phi_stub().quiet) return;
        FILE *f = phi_stub().out();
        fprintf(f, "O %d |", sense); put_lin(f, e); fputc('\n', f);
    }
    void optimize() {
        phi_stub_state &s = phi_stub();
        fprintf(s.out(), "E vars=%ld lin=%ld quad=%ld\n", s.n_vars, s.n_lin, s.n_quad);
        fflush(s.out());
        throw GRBException("recording stub: model dumped, no solver in this image", 10009);
    }
    GRBQuadExpr getObjective() const { return GRBQuadExpr(); }
    GRBVar *getVars() const { return new GRBVar[1]; }
    GRBConstr *getConstrs() const { return new GRBConstr[1]; }
    GRBQConstr *getQConstrs() const { return new GRBQConstr[1]; }
    GRBLinExpr getRow(const GRBConstr &) const { return GRBLinExpr(); }
    GRBQuadExpr getQCRow(const GRBQConstr &) const { return GRBQuadExpr(); }
};

#endif
