// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// A whole CPU planner for comparisons: the host-side local solver of the product (armour_b200/host/local_solver.cpp, the
// stand-in for Ipopt; compiled into liboracle.so from where it lies) driving the ORACLE through the TNLP callbacks, i.e.
// what KPR/armour_main.cu:237-278 does with Ipopt and armtd_NLP, entirely on the CPU.  Used by bench.py as the CPU arm
// of the batched-planning end-to-end number and by tests/test_solver_gpu.py.
#include <cstring>
#include <vector>

#include "../armour_b200/host/local_solver.h"
#include "planner.h"

namespace {
using namespace Ipopt;
struct OracleNLP : public TNLP {
    orc::Problem* P;
    const double* q_des;
    double solution[orc::NF];
    bool feasible = false;
    int first_violation = -1;
    bool armtd = false;  // the ARMTD comparison planner's NLP (KPA/NLPclass.cu) instead of the main planner's
    OracleNLP(orc::Problem* p, const double* qd, bool armtd_ = false) : P(p), q_des(qd), armtd(armtd_) {
        std::memset(solution, 0, sizeof(solution));
    }
    bool get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag, IndexStyleEnum& st) override {
        n = orc::NF;
        m = armtd ? P->armtd_num_constraints() : P->num_constraints();
        nnz_jac_g = m * n;
        nnz_h_lag = 0;
        st = C_STYLE;
        return true;
    }
    bool get_bounds_info(Index n, Number* x_l, Number* x_u, Index, Number* g_l, Number* g_u) override {
        for (Index i = 0; i < n; i++) {  // KPR/NLPclass.cu:105-113
            x_l[i] = -1.0;
            x_u[i] = 1.0;
        }
        if (armtd) P->armtd_bounds(g_l, g_u); else P->bounds(g_l, g_u);
        return true;
    }
    bool get_starting_point(Index n, bool, Number* x, bool, Number*, Number*, Index, bool, Number*) override {
        for (Index i = 0; i < n; i++) x[i] = 0.0;  // KPR/NLPclass.cu:192-197
        return true;
    }
    bool eval_f(Index, const Number* x, bool, Number& obj) override {
        obj = armtd ? P->armtd_cost(q_des, x) : P->cost(q_des, x);
        return true;
    }
    bool eval_grad_f(Index, const Number* x, bool, Number* grad) override {
        if (armtd) P->armtd_cost_grad(q_des, x, grad); else P->cost_grad(q_des, x, grad);
        return true;
    }
    bool eval_g(Index, const Number* x, bool, Index, Number* g) override {
        if (armtd) P->armtd_eval_g(x, g); else P->eval_g(x, g);
        return true;
    }
    bool eval_jac_g(Index n, const Number* x, bool, Index m, Index, Index* iRow, Index* jCol, Number* values) override {
        if (!values) {
            for (Index i = 0; i < m; i++)
                for (Index j = 0; j < n; j++) {
                    iRow[i * n + j] = i;
                    jCol[i * n + j] = j;
                }
        } else {
            if (armtd) P->armtd_eval_jac_g(x, values); else P->eval_jac_g(x, values);
        }
        return true;
    }
    void finalize_solution(SolverReturn, Index n, const Number* x, const Number*, const Number*, Index, const Number* g,
                           const Number*, Number, const IpoptData*, IpoptCalculatedQuantities*) override {
        for (Index i = 0; i < n; i++) solution[i] = x[i];
        feasible = (armtd ? P->armtd_verdict(g, &first_violation) : P->verdict(g, &first_violation)) != 0;  // NLPclass.cu:449-537
    }
};
}  // namespace

extern "C" int orc_solve(void* h, const double* q_des, int max_iter, double max_wall_time, double* k_opt, int* feasible,
                         int* first_violation, int* iterations) {
    orc::Problem* P = static_cast<orc::Problem*>(h);
    OracleNLP nlp(P, q_des);
    LocalSolverOptions opt;
    if (max_iter > 0) opt.max_iter = max_iter;
    opt.max_wall_time = max_wall_time > 0 ? max_wall_time : 1e9;  // no clock by default: results do not depend on the host
    LocalSolverStats st;
    local_solve(nlp, opt, &st);
    for (int i = 0; i < orc::NF; i++) k_opt[i] = nlp.solution[i];
    if (feasible) *feasible = nlp.feasible ? 1 : 0;
    if (first_violation) *first_violation = nlp.first_violation;
    if (iterations) *iterations = st.iterations;
    return 0;
}

// the QP sub-solver of the host solver on its own, for tests/test_active_set_qp.py
extern "C" int orc_qp(int n, double h, const double* c, int nrows, const double* A, const double* b, double* d) {
    return local_qp(n, h, c, nrows, A, b, d, 200);
}

// the same for the ARMTD comparison planner (KPA/armtd_main.cu:163-205 with the local solver in Ipopt's place; tol as
// KPA/Parameters.h:43); fills the sliced link centres at the solution like the reference's last eval_g
extern "C" int orc_armtd_solve(void* h, const double* q_des, int max_iter, double tol, double* k_opt, int* feasible,
                               int* first_violation, int* iterations) {
    orc::Problem* P = static_cast<orc::Problem*>(h);
    OracleNLP nlp(P, q_des, true);
    LocalSolverOptions opt;
    if (max_iter > 0) opt.max_iter = max_iter;
    if (tol > 0) opt.tol = tol;
    opt.max_wall_time = 1e9;
    LocalSolverStats st;
    local_solve(nlp, opt, &st);
    for (int i = 0; i < orc::NF; i++) k_opt[i] = nlp.solution[i];
    if (feasible) *feasible = nlp.feasible ? 1 : 0;
    if (first_violation) *first_violation = nlp.first_violation;
    if (iterations) *iterations = st.iterations;
    std::vector<double> g(P->armtd_num_constraints());
    P->armtd_eval_g(nlp.solution, g.data());
    return 0;
}
