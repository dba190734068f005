#include "local_solver.h"

#include "active_set_qp.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <vector>

using namespace Ipopt;

namespace {

constexpr int NMAX = 16;  // most optimisation variables (the planner has 7)

struct Lin {  // one linearised inequality  a . d <= b
    double a[NMAX];
    double b;
    double scale;  // 1 / |a|
};

// min 1/2 h |d|^2 + c . d  s.t.  a_i . d <= b_i, exactly (active_set_qp.h).  Returns the primal step; iters / drops
// for the statistics.
void solve_qp(int n, double h, const double* c, const std::vector<Lin>& rows, double* d, int max_outer, int* iters) {
    asqp::State<NMAX> S;
    asqp::init(S, n, h, c);
    int it = 0;
    for (; it < max_outer; it++) {
        double best = asqp::VIOLATION_TOL;
        int p = -1;
        for (size_t i = 0; i < rows.size(); i++) {
            const double v = asqp::row_violation(rows[i].a, rows[i].b, rows[i].scale, S.d, n);
            if (v > best && !asqp::is_excluded(S, int(i))) {
                best = v;
                p = int(i);
            }
        }
        if (p < 0) break;
        if (asqp::add_row(S, p, rows[p].a, rows[p].b) == 2) break;
    }
    for (int j = 0; j < n; j++) d[j] = S.d[j];
    if (iters) *iters = it;
}

}  // namespace

int local_qp(int n, double h, const double* c, int nrows, const double* A, const double* b, double* d, int max_outer) {
    if (n < 1 || n > NMAX || !(h > 0)) return -1;
    std::vector<Lin> rows(nrows);
    for (int i = 0; i < nrows; i++) {
        for (int j = 0; j < n; j++) rows[i].a[j] = A[size_t(i) * n + j];
        rows[i].b = b[i];
        rows[i].scale = asqp::row_scale(rows[i].a, n);
    }
    int it = 0;
    solve_qp(n, h, c, rows, d, max_outer, &it);
    return it;
}

SolverReturn local_solve(TNLP& nlp, const LocalSolverOptions& opt, LocalSolverStats* stats) {
    const auto t0 = std::chrono::steady_clock::now();
    auto elapsed = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    Index n = 0, m = 0, nnz = 0, nnzh = 0;
    TNLP::IndexStyleEnum style;
    nlp.get_nlp_info(n, m, nnz, nnzh, style);
    std::vector<Number> xl(n), xu(n), gl(m), gu(m), x(n), g(m), J(size_t(m) * n), gf(n), xt(n), gt(m), best(n), gbest(m);
    nlp.get_bounds_info(n, xl.data(), xu.data(), m, gl.data(), gu.data());
    nlp.get_starting_point(n, true, x.data(), false, nullptr, nullptr, m, false, nullptr);
    LocalSolverStats st;

    // violation with the verdict's tolerances folded in: <= 0 means "finalize_solution would accept"
    const Index n_torque = 7 * ((m > 28) ? 1 : 0);  // row classes are recognised from the bounds, not from indices
    (void)n_torque;
    auto row_tol = [&](Index i) {
        if (gl[i] <= -1e18) return opt.collision_tol;                    // one-sided rows: collision
        if (i < m - 4 * n) return opt.torque_tol;                        // two-sided rows before the last 4n: torque
        return 0.0;                                                      // joint position / velocity limits
    };
    auto violation = [&](const std::vector<Number>& gv) {
        double v = -1e300;
        for (Index i = 0; i < m; i++) {
            const double tol = row_tol(i);
            v = std::max(v, gv[i] - gu[i] - tol);
            if (gl[i] > -1e18) v = std::max(v, gl[i] - gv[i] - tol);
        }
        return v;
    };

    Number f = 0;
    nlp.eval_f(n, x.data(), true, f);
    nlp.eval_g(n, x.data(), true, m, g.data());
    st.evals++;
    double viol = violation(g);
    bool have_best = viol <= 0;
    double fbest = f;
    if (have_best) {
        best = x;
        gbest = g;
    }
    double delta = 0.5;
    SolverReturn status = MAXITER_EXCEEDED;
    for (int it = 0; it < opt.max_iter; it++) {
        if (elapsed() > opt.max_wall_time) {
            status = CPUTIME_EXCEEDED;
            break;
        }
        st.iterations = it + 1;
        nlp.eval_grad_f(n, x.data(), false, gf.data());
        nlp.eval_jac_g(n, x.data(), false, m, nnz, nullptr, nullptr, J.data());
        // curvature of the cost along the gradient from one extra evaluation (exact for the planner's quadratic cost)
        double h = 1.0;
        {
            double gn2 = 0;
            for (Index j = 0; j < n; j++) gn2 += gf[j] * gf[j];
            if (gn2 > 0) {
                const double eps = 1e-3 / std::sqrt(gn2);
                for (Index j = 0; j < n; j++) xt[j] = x[j] - eps * gf[j];
                Number f2 = 0;
                nlp.eval_f(n, xt.data(), true, f2);
                const double curv = 2.0 * (f2 - f + eps * gn2) / (eps * eps * gn2);
                if (curv > 1e-8) h = curv;
            }
        }
        // rows that can be reached inside the trust region, linearised
        std::vector<Lin> rows;
        auto push = [&](const Number* a, double sign, double b) {
            Lin r;
            double l1 = 0;
            for (Index j = 0; j < n; j++) {
                r.a[j] = sign * a[j];
                l1 += std::fabs(r.a[j]);
            }
            if (b > l1 * delta) return;  // cannot become active within |d|_inf <= delta
            r.b = b;
            r.scale = asqp::row_scale(r.a, n);
            rows.push_back(r);
        };
        for (Index i = 0; i < m; i++) {
            const Number* a = &J[size_t(i) * n];
            const double tol = 0.5 * row_tol(i);  // aim inside the acceptance band
            push(a, 1.0, gu[i] + tol - g[i]);
            if (gl[i] > -1e18) push(a, -1.0, g[i] - (gl[i] - tol));
        }
        std::vector<Number> e(n, 0.0);
        for (Index j = 0; j < n; j++) {
            e.assign(n, 0.0);
            e[j] = 1.0;
            push(e.data(), 1.0, std::min(delta, xu[j] - x[j]));
            push(e.data(), -1.0, std::min(delta, x[j] - xl[j]));
        }
        std::vector<double> d(n);
        int qp_it = 0;
        solve_qp(n, h, gf.data(), rows, d.data(), opt.qp_sweeps, &qp_it);
        st.qp_iterations += qp_it;
        double dn = 0;
        for (Index j = 0; j < n; j++) {
            d[j] = std::max(-delta, std::min(delta, d[j]));
            xt[j] = std::max(xl[j], std::min(xu[j], x[j] + d[j]));
            dn = std::max(dn, std::fabs(xt[j] - x[j]));
        }
        if (dn < opt.tol) {
            status = (viol <= 0) ? SUCCESS : LOCAL_INFEASIBILITY;
            break;
        }
        Number ft = 0;
        nlp.eval_f(n, xt.data(), true, ft);
        nlp.eval_g(n, xt.data(), true, m, gt.data());
        st.evals++;
        const double vt = violation(gt);
        // filter-style acceptance: feasible points must lower the cost, infeasible ones must lower the violation
        const bool accept = (vt <= 0 && (viol > 0 || ft < f - 1e-12)) || (vt > 0 && viol > 0 && vt < viol - 1e-12);
        if (accept) {
            x = xt;
            g = gt;
            f = ft;
            viol = vt;
            delta = std::min(1.0, delta * 1.5);
            if (viol <= 0 && (!have_best || f < fbest)) {
                have_best = true;
                fbest = f;
                best = x;
                gbest = g;
            }
        } else {
            delta *= 0.4;
            if (delta < opt.tol) {
                status = (viol <= 0) ? STOP_AT_TINY_STEP : LOCAL_INFEASIBILITY;
                break;
            }
        }
    }
    if (have_best) {
        x = best;
        g = gbest;
        f = fbest;
        viol = violation(g);
    }
    st.final_violation = viol;
    st.final_cost = f;
    st.seconds = elapsed();
    if (stats) *stats = st;
    nlp.finalize_solution(status, n, x.data(), nullptr, nullptr, m, g.data(), nullptr, f, nullptr, nullptr);
    return status;
}
