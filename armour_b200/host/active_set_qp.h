// Exact solver of the trust-region QP of one SQP step, shared by the host solver (local_solver.cpp) and the
// device solver (csrc/k4_solver.cuh): the same statements in the same order on both sides, so both produce the
// same iterates.
//
//     min  1/2 h |d|^2 + c . d     s.t.   a_i . d <= b_i ,  i = 0 .. rows-1          (h > 0, d in R^n, n <= NMAX)
//
// Dual active-set method (Goldfarb & Idnani 1983) specialised to a spherical Hessian: start from the unconstrained
// minimiser d = -c / h, repeatedly take the most violated row (scaled by 1 / |a_i|; ties: lowest index) into the
// active set, dropping rows whose multiplier would turn negative on the way.  Every iterate is the exact minimiser
// over its active rows, the dual objective grows strictly, so the method ends after finitely many steps at the QP's
// optimum — no sweep cap to run into (the round-1 Hildreth iteration ended at its 200-sweep cap on most QPs).
// A row that contradicts the rows already active (no primal or dual step exists) is an infeasible linearisation:
// it is put on a skip list and the others go on, so the step still satisfies a greedy maximal subset.
//
// The caller owns the search for the most violated row (serial on the host, one CTA on the device); this header
// holds the state and the algebra of adding one row, which is O(n^3) with n = 7.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define ASQP_HD __host__ __device__ inline
#else
#define ASQP_HD inline
#endif

namespace asqp {

constexpr double VIOLATION_TOL = 1e-11;   // scaled violation (a.d - b) / |a| above which a row counts as violated
constexpr double RANK_TOL = 1e-18;        // |component of a_p outside the active span|^2 / |a_p|^2 below this: dependent
constexpr int MAX_SKIP = 8;               // contradicting rows remembered per QP
constexpr int MAX_OUTER = 200;            // rows taken in per QP (a guard; the method is finite)

template <int NMAX>
struct State {
    int n, nact, nskip, drops;
    int nq;                    // leading active rows whose columns of Q and R are valid
    double h;
    double d[NMAX];            // primal iterate: the minimiser over the active rows
    double u[NMAX];            // multipliers of the active rows (>= 0)
    int idx[NMAX];             // their row indices
    double A[NMAX][NMAX];      // their normals
    double Q[NMAX][NMAX];      // orthonormal basis of the normals (modified Gram-Schmidt, N = Q R), kept between rows:
    double R[NMAX][NMAX];      // a new row appends one column, a dropped row invalidates the columns from its position on
    int skip[MAX_SKIP];
};

// 1 / |a| of a row (0 for a zero row, which is then never selected)
ASQP_HD double row_scale(const double* a, int n) {
    double aa = 0;
    for (int j = 0; j < n; j++) aa += a[j] * a[j];
    return aa > 0 ? 1.0 / std::sqrt(aa) : 0.0;
}

// scaled violation of a row at d
ASQP_HD double row_violation(const double* a, double b, double scale, const double* d, int n) {
    double s = -b;
    for (int j = 0; j < n; j++) s += a[j] * d[j];
    return s * scale;
}

template <int NMAX>
ASQP_HD void init(State<NMAX>& S, int n, double h, const double* c) {
    S.n = n;
    S.nact = 0;
    S.nskip = 0;
    S.drops = 0;
    S.nq = 0;
    S.h = h;
    for (int j = 0; j < n; j++) S.d[j] = -c[j] / h;
}

template <int NMAX>
ASQP_HD bool is_excluded(const State<NMAX>& S, int i) {
    for (int q = 0; q < S.nact; q++)
        if (S.idx[q] == i) return true;
    for (int q = 0; q < S.nskip; q++)
        if (S.skip[q] == i) return true;
    return false;
}

// Take the violated row p (normal a, right side b) into the active set.  Returns 0 when it was added (d now satisfies
// it with equality), 1 when it contradicts the active rows and went on the skip list, 2 when the skip list is full
// (the caller stops; d is the last consistent iterate).
template <int NMAX>
ASQP_HD int add_row(State<NMAX>& S, int p, const double* a, double b) {
    const int n = S.n;
    double up = 0.0;
    double ap2 = 0;
    for (int j = 0; j < n; j++) ap2 += a[j] * a[j];
    for (int guard = 0; guard < 2 * NMAX + 2; guard++) {
        const int na = S.nact;
        // orthonormal basis of the active normals by modified Gram-Schmidt, N = Q R: the columns that are not valid yet
        double y[NMAX], r[NMAX], w[NMAX];
        for (int k = S.nq; k < na; k++) {
            for (int j = 0; j < n; j++) S.Q[k][j] = S.A[k][j];
            for (int i = 0; i < k; i++) {
                double dot = 0;
                for (int j = 0; j < n; j++) dot += S.Q[i][j] * S.Q[k][j];
                S.R[i][k] = dot;
                for (int j = 0; j < n; j++) S.Q[k][j] -= dot * S.Q[i][j];
            }
            double nn = 0;
            for (int j = 0; j < n; j++) nn += S.Q[k][j] * S.Q[k][j];
            const double nr = std::sqrt(nn);
            S.R[k][k] = nr;
            const double inv = nr > 0 ? 1.0 / nr : 0.0;
            for (int j = 0; j < n; j++) S.Q[k][j] *= inv;
        }
        S.nq = na;
        // w = part of a outside the active span (the primal step direction, times h), y = Q^T a
        for (int j = 0; j < n; j++) w[j] = a[j];
        for (int k = 0; k < na; k++) {
            double dot = 0;
            for (int j = 0; j < n; j++) dot += S.Q[k][j] * w[j];
            y[k] = dot;
            for (int j = 0; j < n; j++) w[j] -= dot * S.Q[k][j];
        }
        // r = R^-1 y: how the active multipliers give way per unit of the new one
        for (int k = na - 1; k >= 0; k--) {
            double s = y[k];
            for (int i = k + 1; i < na; i++) s -= S.R[k][i] * r[i];
            r[k] = S.R[k][k] > 0 ? s / S.R[k][k] : 0.0;
        }
        double zz = 0;
        for (int j = 0; j < n; j++) zz += w[j] * w[j];
        const bool primal = na < n && zz > RANK_TOL * ap2;
        // longest dual step that keeps the active multipliers non-negative
        double t1 = HUGE_VAL;
        int jdrop = -1;
        for (int k = 0; k < na; k++) {
            if (r[k] > 0) {
                const double tt = S.u[k] / r[k];
                if (tt < t1) {
                    t1 = tt;
                    jdrop = k;
                }
            }
        }
        // step that makes row p active
        double t2 = HUGE_VAL;
        if (primal) {
            double s = -b;
            for (int j = 0; j < n; j++) s += a[j] * S.d[j];
            t2 = s > 0 ? s * S.h / zz : 0.0;
        }
        if (jdrop < 0 && !primal) {  // no step at all: row p contradicts the active rows
            // give back the dual steps taken for p so far?  They moved d only along directions that keep the active
            // rows tight and reduce p's violation, and kept u >= 0: the iterate is the minimiser over the (reduced)
            // active set plus a multiple of a_p — still a descent compromise; keep it.
            if (S.nskip >= MAX_SKIP) return 2;
            S.skip[S.nskip++] = p;
            return 1;
        }
        if (t2 <= t1) {  // full step: p becomes active
            const double sc = t2 / S.h;
            for (int j = 0; j < n; j++) S.d[j] -= sc * w[j];
            for (int k = 0; k < na; k++) {
                const double nu = S.u[k] - t2 * r[k];
                S.u[k] = nu > 0 ? nu : 0.0;
            }
            up += t2;
            for (int j = 0; j < n; j++) S.A[na][j] = a[j];
            S.u[na] = up;
            S.idx[na] = p;
            S.nact = na + 1;
            // the new column of Q R is what was just computed: w is a orthogonalised against the columns before it
            {
                for (int i = 0; i < na; i++) S.R[i][na] = y[i];
                const double nr = std::sqrt(zz);
                S.R[na][na] = nr;
                const double inv = nr > 0 ? 1.0 / nr : 0.0;
                for (int j = 0; j < n; j++) S.Q[na][j] = w[j] * inv;
                S.nq = na + 1;
            }
            return 0;
        }
        // partial step: multiplier jdrop reaches zero first; drop that row and try again
        if (primal) {
            const double sc = t1 / S.h;
            for (int j = 0; j < n; j++) S.d[j] -= sc * w[j];
        }
        for (int k = 0; k < na; k++) {
            const double nu = S.u[k] - t1 * r[k];
            S.u[k] = nu > 0 ? nu : 0.0;
        }
        up += t1;
        for (int k = jdrop; k + 1 < na; k++) {
            for (int j = 0; j < n; j++) S.A[k][j] = S.A[k + 1][j];
            S.u[k] = S.u[k + 1];
            S.idx[k] = S.idx[k + 1];
        }
        S.nact = na - 1;
        if (S.nq > jdrop) S.nq = jdrop;  // the columns from the dropped position on are recomputed at the next pass
        S.drops++;
    }
    if (S.nskip >= MAX_SKIP) return 2;
    S.skip[S.nskip++] = p;
    return 1;
}

}  // namespace asqp
