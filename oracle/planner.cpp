// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned against the reference's own sources: reach-set build, slices and Bezier rows through oracle/_ref/libarmour_ref.so (tests/test_oracle_pinned.py); collision rows, bounds, cost and verdict through the reference's CUDA kernels and armtd_NLP compiled by nvcc (oracle/_ref/libarmour_ref_cuda.so, run on a B200, frozen as tests/golden/refcuda/, tests/test_refcuda_golden.py).
// See planner.h.
#include "planner.h"

#include <omp.h>

#include <chrono>
#include <cmath>
#include <cstring>

namespace orc {

Problem::Problem(int model_id, const PlannerParams& p) : model(make_robot_model(model_id)), params(p) {
    T = params.num_time_steps;
    NJ = model.num_joints;
}

void Problem::build(const double* q0, const double* qd0, const double* qdd0, const double* obs, int nobs,
                    int nthreads) {
    if (nobs > params.max_obstacles || nobs < 0) throw -1;  // CollisionChecking.cu:10-13
    O = nobs;
    obstacles.assign(obs, obs + size_t(nobs) * 12);
    const double thr = params.simplify_threshold;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    stats = Stats();
    const auto t0 = std::chrono::steady_clock::now();

    tls_threshold() = thr;
    traj.reset(new BezierCurve(&model, &params, q0, qd0, qdd0));
    BezierCurve& tr = *traj;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int t = 0; t < T; t++) {  // armour_main.cu:98-102
        tls_threshold() = thr;
        tls_stats() = Stats();
        tr.makePolyZono(t);
#pragma omp critical
        stats.add(tls_stats());
    }

    tls_stats() = Stats();
    kd.reset(new KinematicsDynamics(traj.get()));
    stats.add(tls_stats());
    KinematicsDynamics& K = *kd;
    link_gens.assign(size_t(T) * NJ * 18, 0.0);
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
    for (int t = 0; t < T; t++) {  // armour_main.cu:116-142
        tls_threshold() = thr;
        tls_stats() = Stats();
        K.fk(t);
        for (int i = 0; i < NJ; i++) K.links[i * T + t].reduce_link_PZ(&link_gens[size_t(t * NJ + i) * 18]);
        K.rnea_nominal(t);
        K.rnea_interval(t);
        for (int i = 0; i < NF; i++) K.u_nom_int[i * T + t] = K.u_nom_int[i * T + t] - K.u_nom[i * T + t];
        for (int i = 0; i < NF; i++) K.u_nom[i * T + t].reduce();
#pragma omp critical
        stats.add(tls_stats());
    }

    // robust input bound, armour_main.cu:172-201
    torque_radius.assign(size_t(NF) * T, 0.0);
    for (int t = 0; t < T; t++) {
        Interval rho(0.0);
        for (int i = 0; i < NF; i++) {
            Interval w;
            K.u_nom_int[i * T + t].to_interval(&w);
            rho = rho + w * w;
            torque_radius[i * T + t] =
                model.alpha * (model.M_max - model.M_min) * model.eps + 0.5 * std::max(std::fabs(w.lo), std::fabs(w.hi));
        }
        rho = sqrt(rho);
        for (int i = 0; i < NF; i++) torque_radius[i * T + t] += 0.5 * rho.hi;
        for (int i = 0; i < NF; i++) torque_radius[i * T + t] += K.u_nom[i * T + t].indep[0];
        for (int i = 0; i < NF; i++) torque_radius[i * T + t] += model.friction[i];
    }

    init_hyperplanes();
    build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

    link_sliced_center.assign(size_t(T) * NJ * 3, 0.0);
    dk_link_sliced_center.assign(size_t(T) * NJ * NF * 3, 0.0);
}

// bufferObstaclesKernel + polytope_PH, CollisionChecking.cu:136-228 (pair order :26-39)
void Problem::init_hyperplanes() {
    const size_t n = size_t(T) * NJ * O * kComb;
    A.assign(n * 3, 0.0);
    d.assign(n, 0.0);
    delta.assign(n, 0.0);
    int combA[kComb], combB[kComb];
    {
        int a = 0, b = 1;
        for (int i = 0; i < kComb; i++) {
            combA[i] = a;
            combB[i] = b;
            if (b < kBufGen - 1) {
                b++;
            } else {
                a++;
                b = a + 1;
            }
        }
    }
    for (int t = 0; t < T; t++)
        for (int l = 0; l < NJ; l++) {
            const double* LG = &link_gens[size_t(t * NJ + l) * 18];
            for (int o = 0; o < O; o++) {
                double G[kBufGen][3], c[3];
                for (int p = 0; p < 3; p++) {
                    c[p] = obstacles[o * 12 + p];
                    for (int i = 0; i < 3; i++) G[i][p] = obstacles[(o * 4 + i + 1) * 3 + p];
                    for (int i = 0; i < 6; i++) G[i + 3][p] = LG[p + i * 3];
                }
                for (int p = 0; p < kComb; p++) {
                    const double* ga = G[combA[p]];
                    const double* gb = G[combB[p]];
                    double cr[3] = {ga[1] * gb[2] - ga[2] * gb[1], ga[2] * gb[0] - ga[0] * gb[2],
                                    ga[0] * gb[1] - ga[1] * gb[0]};
                    const double nrm = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
                    double C[3] = {0, 0, 0};
                    if (nrm > 0) {
                        for (int e = 0; e < 3; e++) C[e] = cr[e] / nrm;
                    }
                    const size_t idx = (size_t(t * NJ + l) * O + o) * kComb + p;
                    for (int e = 0; e < 3; e++) A[idx * 3 + e] = C[e];
                    d[idx] = C[0] * c[0] + C[1] * c[1] + C[2] * c[2];
                    double dl = 0.0;
                    for (int j = 0; j < kBufGen; j++) dl += std::fabs(C[0] * G[j][0] + C[1] * G[j][1] + C[2] * G[j][2]);
                    delta[idx] = dl;
                }
            }
        }
}

// linkFRSConstraints + checkCollisionKernel, CollisionChecking.cu:90-134,230-299
// link_c: [l][t][o], grad_link_c: [l][t][o][NF]
void Problem::link_constraints(bool with_grad, double* link_c, double* grad_link_c) {
    for (int l = 0; l < NJ; l++)
        for (int t = 0; t < T; t++) {
            const double* ctr = &link_sliced_center[size_t(t * NJ + l) * 3];
            for (int o = 0; o < O; o++) {
                const size_t base = (size_t(t * NJ + l) * O + o) * kComb;
                double max_elt = -100000000;
                int max_id = 0;
                bool neg = false;
                for (int p = 0; p < kComb; p++) {
                    const double* a = &A[(base + p) * 3];
                    double pos_res, neg_res;
                    if (std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]) > 0) {
                        const double dot = a[0] * ctr[0] + a[1] * ctr[1] + a[2] * ctr[2];
                        pos_res = dot - (d[base + p] + delta[base + p]);
                        neg_res = -dot - (-d[base + p] + delta[base + p]);
                    } else {
                        pos_res = -100000000;
                        neg_res = -100000000;
                    }
                    if (pos_res > max_elt) {
                        max_elt = pos_res;
                        max_id = p;
                        neg = false;
                    }
                    if (neg_res > max_elt) {
                        max_elt = neg_res;
                        max_id = p;
                        neg = true;
                    }
                }
                const size_t row = (size_t(l) * T + t) * O + o;
                if (link_c) link_c[row] = -max_elt;
                if (with_grad) {
                    const double* a = &A[(base + max_id) * 3];
                    for (int v = 0; v < NF; v++) {
                        const double* dk = &dk_link_sliced_center[(size_t(t * NJ + l) * NF + v) * 3];
                        const double dot = a[0] * dk[0] + a[1] * dk[1] + a[2] * dk[2];
                        grad_link_c[row * NF + v] = neg ? dot : -dot;
                    }
                }
            }
        }
}

void Problem::eval_g(const double* k, double* g) {  // NLPclass.cu:272-324 (input constraints on)
    KinematicsDynamics& K = *kd;
#pragma omp parallel for schedule(dynamic)
    for (int t = 0; t < T; t++) {
        for (int j = 0; j < NF; j++) {
            double lo, hi;
            K.u_nom[j * T + t].slice(k, &lo, &hi);
            g[t * NF + j] = (lo + hi) * 0.5;
        }
        for (int l = 0; l < NJ; l++) {
            double lo[3], hi[3];
            K.links[l * T + t].slice(k, lo, hi);
            for (int e = 0; e < 3; e++) link_sliced_center[size_t(t * NJ + l) * 3 + e] = (lo[e] + hi[e]) * 0.5;
        }
    }
    link_constraints(false, g + T * NF, nullptr);
    traj->jointPositionExtremum(g + T * NF + T * NJ * O, k);
    traj->jointVelocityExtremum(g + T * NF + T * NJ * O + NF * 2, k);
}

void Problem::eval_jac_g(const double* k, double* values) {  // NLPclass.cu:330-396
    KinematicsDynamics& K = *kd;
#pragma omp parallel for schedule(dynamic)
    for (int t = 0; t < T; t++) {
        for (int j = 0; j < NF; j++) K.u_nom[j * T + t].slice_gradient(k, values + size_t(t * NF + j) * NF);
        for (int l = 0; l < NJ; l++) {
            double lo[3], hi[3];
            K.links[l * T + t].slice(k, lo, hi);
            for (int e = 0; e < 3; e++) link_sliced_center[size_t(t * NJ + l) * 3 + e] = (lo[e] + hi[e]) * 0.5;
            K.links[l * T + t].slice_gradient(k, &dk_link_sliced_center[size_t(t * NJ + l) * NF * 3]);
        }
    }
    link_constraints(true, nullptr, values + size_t(T) * NF * NF);
    traj->jointPositionExtremumGradient(values + size_t(T * NF + T * NJ * O) * NF, k);
    traj->jointVelocityExtremumGradient(values + size_t(T * NF + T * NJ * O + NF * 2) * NF, k);
}

void Problem::bounds(double* g_l, double* g_u) const {  // NLPclass.cu:116-165
    int offset = 0;
    for (int i = 0; i < T; i++)
        for (int j = 0; j < NF; j++) {
            g_l[i * NF + j] = -model.torque_limits[j] + torque_radius[j * T + i];
            g_u[i * NF + j] = model.torque_limits[j] - torque_radius[j * T + i];
        }
    offset += NF * T;
    for (int i = offset; i < offset + T * NJ * O; i++) {
        g_l[i] = -1e19;
        g_u[i] = 0;
    }
    offset += T * NJ * O;
    for (int rep = 0; rep < 2; rep++) {
        for (int i = 0; i < NF; i++) {
            g_l[offset + i] = model.state_limits_lb[i] + model.qe;
            g_u[offset + i] = model.state_limits_ub[i] - model.qe;
        }
        offset += NF;
    }
    for (int rep = 0; rep < 2; rep++) {
        for (int i = 0; i < NF; i++) {
            g_l[offset + i] = -model.speed_limits[i] + model.qde;
            g_u[offset + i] = model.speed_limits[i] - model.qde;
        }
        offset += NF;
    }
}

int Problem::verdict(const double* g, int* first) const {  // NLPclass.cu:449-537
    auto fail = [&](int row) {
        if (first) *first = row;
        return 0;
    };
    int offset = 0;
    for (int i = 0; i < T; i++)
        for (int j = 0; j < NF; j++) {
            const double v = g[i * NF + j];
            if (v < -model.torque_limits[j] + torque_radius[j * T + i] - params.torque_violation_threshold ||
                v > model.torque_limits[j] - torque_radius[j * T + i] + params.torque_violation_threshold)
                return fail(i * NF + j);
        }
    offset += NF * T;
    for (int i = 0; i < NJ; i++)
        for (int j = 0; j < T; j++)
            for (int h = 0; h < O; h++)
                if (g[(i * T + j) * O + h + offset] > params.collision_violation_threshold)
                    return fail((i * T + j) * O + h + offset);
    offset += NJ * T * O;
    for (int rep = 0; rep < 2; rep++) {
        for (int i = offset; i < offset + NF; i++)
            if (g[i] < model.state_limits_lb[i - offset] + model.qe || g[i] > model.state_limits_ub[i - offset] - model.qe)
                return fail(i);
        offset += NF;
    }
    for (int rep = 0; rep < 2; rep++) {
        for (int i = offset; i < offset + NF; i++)
            if (g[i] < -model.speed_limits[i - offset] + model.qde || g[i] > model.speed_limits[i - offset] - model.qde)
                return fail(i);
        offset += NF;
    }
    if (first) *first = -1;
    return 1;
}

static double wrap_to_pi(double angle) {  // NLPclass.cu:6-15
    double w = angle;
    while (w < -M_PI) w += 2 * M_PI;
    while (w > M_PI) w -= 2 * M_PI;
    return w;
}

double Problem::cost(const double* q_des, const double* k) const {  // NLPclass.cu:207-236
    double q_plan[NF];
    for (int i = 0; i < NF; i++)
        q_plan[i] = q_des_func(traj->q0[i], traj->Tqd0[i], traj->TTqdd0[i], params.k_range[i] * k[i], params.t_plan);
    double obj = std::pow(wrap_to_pi(q_des[0] - q_plan[0]), 2) + std::pow(wrap_to_pi(q_des[2] - q_plan[2]), 2) +
                 std::pow(wrap_to_pi(q_des[4] - q_plan[4]), 2) + std::pow(wrap_to_pi(q_des[6] - q_plan[6]), 2) +
                 std::pow(q_des[1] - q_plan[1], 2) + std::pow(q_des[3] - q_plan[3], 2) +
                 std::pow(q_des[5] - q_plan[5], 2);
    return obj * params.cost_scale;
}

void Problem::cost_grad(const double* q_des, const double* k, double* grad) const {  // NLPclass.cu:241-268
    const double tp = params.t_plan;
    for (int i = 0; i < NF; i++) {
        const double q_plan = q_des_func(traj->q0[i], traj->Tqd0[i], traj->TTqdd0[i], params.k_range[i] * k[i], tp);
        const double dk = std::pow(tp, 3) * (6 * std::pow(tp, 2) - 15 * tp + 10) * params.k_range[i];
        if (i % 2 == 0)
            grad[i] = (2 * wrap_to_pi(q_plan - q_des[i]) * dk);
        else
            grad[i] = (2 * (q_plan - q_des[i]) * dk);
        grad[i] *= params.cost_scale;
    }
}

}  // namespace orc
