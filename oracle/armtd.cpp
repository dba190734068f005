// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned against the reference's own KPA sources compiled by nvcc
// (oracle/_ref/libarmour_ref_armtd.so, run on a B200, frozen as tests/golden/armtd/reference.npz; tests/test_armtd_oracle.py).
//
// CPU restatement of the ARMTD comparison planner (reference directory
// kinova_planner_realtime_armtd_comparison = "KPA"): what differs from the main planner is the trajectory class
// (KPA/Trajectory.cu: ConstantAccelerationCurve — rotation PZs from an OFFLINE joint reachable set handed in by the caller,
// closed-form joint position / velocity extrema) and the NLP (KPA/NLPclass.cu: no torque rows, other cost).  The PZ arithmetic,
// the forward kinematics, reduce_link_PZ and the collision rows are the main planner's (KPA's copies of those files are
// identical up to messages) and are reused from pz.cpp / dynamics.cpp / planner.cpp.
#include <omp.h>

#include <cmath>
#include <cstring>

#include "planner.h"

namespace orc {

void Problem::build_armtd(const double* q0, const double* qd0, const double* jrs, const double* k_range, const double* obs,
                          int nobs, int nthreads) {
    if (nobs > params.max_obstacles || nobs < 0) throw -1;  // KPA/armtd_main.cu:90-95
    O = nobs;
    obstacles.assign(obs, obs + size_t(nobs) * 12);
    const double thr = params.simplify_threshold;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    for (int i = 0; i < NF; i++) {
        a_q0[i] = q0[i];
        a_qd0[i] = qd0[i];
        a_k_range[i] = k_range[i];
    }
    tls_threshold() = thr;
    const double zero[NF] = {0, 0, 0, 0, 0, 0, 0};
    traj.reset(new BezierCurve(&model, &params, q0, qd0, zero));  // container of the rotation PZs only
    BezierCurve& tr = *traj;
    const RobotModel& m = model;
    auto at = [&](int a, int i, int t) { return jrs[(size_t(a) * NF + i) * T + t]; };
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int t = 0; t < T; t++) {  // ConstantAccelerationCurve::makePolyZono, KPA/Trajectory.cu:29-86
        tls_threshold() = thr;
        for (int i = 0; i < NF; i++) {
            const double cos_q0 = std::cos(q0[i]), sin_q0 = std::sin(q0[i]);
            const double cos_c = cos_q0 * at(0, i, t) - sin_q0 * at(3, i, t);
            double cos_coeff[2];
            cos_coeff[0] = cos_q0 * at(1, i, t) - sin_q0 * at(4, i, t);
            cos_coeff[1] = std::fabs(cos_q0) * at(2, i, t) + std::fabs(sin_q0) * at(5, i, t);
            cos_coeff[1] *= 4.0;
            const uint64_t cos_hash[2] = {var_hash(i), var_hash(i + NF * 4)};
            const double sin_c = cos_q0 * at(3, i, t) + sin_q0 * at(0, i, t);
            double sin_coeff[2];
            sin_coeff[0] = cos_q0 * at(4, i, t) + sin_q0 * at(1, i, t);
            sin_coeff[1] = std::fabs(cos_q0) * at(5, i, t) + std::fabs(sin_q0) * at(2, i, t);
            sin_coeff[1] *= 4.0;
            const uint64_t sin_hash[2] = {var_hash(i), var_hash(i + NF * 5)};
            PZ Ri = PZ::rpy(m.rots[i * 3], m.rots[i * 3 + 1], m.rots[i * 3 + 2]);
            if (m.axes[i] != 0) Ri = Ri * PZ::rotation(cos_c, cos_coeff, cos_hash, 2, sin_c, sin_coeff, sin_hash, 2, m.axes[i]);
            tr.R[i * T + t] = Ri;
            tr.R_t[i * T + t] = Ri.transpose();
        }
        for (int i = NF; i < m.num_joints; i++) {
            tr.R[i * T + t] = PZ::rpy(m.rots[i * 3], m.rots[i * 3 + 1], m.rots[i * 3 + 2]);
            tr.R_t[i * T + t] = tr.R[i * T + t].transpose();
        }
        tr.R[m.num_joints * T + t] = PZ::rpy(0, 0, 0);
    }
    kd.reset(new KinematicsDynamics(traj.get()));
    KinematicsDynamics& K = *kd;
    link_gens.assign(size_t(T) * NJ * 18, 0.0);
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
    for (int t = 0; t < T; t++) {  // armtd_main.cu:140-149
        tls_threshold() = thr;
        K.fk(t);
        for (int i = 0; i < NJ; i++) K.links[i * T + t].reduce_link_PZ(&link_gens[size_t(t * NJ + i) * 18]);
    }
    init_hyperplanes();
    link_sliced_center.assign(size_t(T) * NJ * 3, 0.0);
    dk_link_sliced_center.assign(size_t(T) * NJ * NF * 3, 0.0);
}

void Problem::armtd_state_extremum(const double* k, double* ext, double* grad) const {  // KPA/Trajectory.cu:88-384
    const double t_move = 0.5, t_total = 1.0, t_to_stop = t_total - t_move;
    for (int i = 0; i < NF; i++) {
        const double q0 = a_q0[i], qd0 = a_qd0[i];
        const double k_actual = a_k_range[i] * k[i];
        const double q_peak = q0 + qd0 * t_move + k_actual * t_move * t_move * 0.5;
        const double q_dot_peak = qd0 + k_actual * t_move;
        const double q_ddot_to_stop = -q_dot_peak / t_to_stop;
        const double q_stop = q_peak + q_dot_peak * t_to_stop + 0.5 * q_ddot_to_stop * t_to_stop * t_to_stop;
        const double t_mm = -qd0 / k_actual;  // time of the interior extremum of the first phase (inf / nan for k = 0: not taken)
        double q_lo, q_hi, g_lo, g_hi;        // end points of the first phase, ordered
        if (q_peak >= q0) {
            q_lo = q0; q_hi = q_peak; g_lo = 0; g_hi = 0.5 * t_move * t_move;
        } else {
            q_lo = q_peak; q_hi = q0; g_lo = 0.5 * t_move * t_move; g_hi = 0;
        }
        double q_min_p, q_max_p, gq_min_p, gq_max_p;
        if (t_mm > 0 && t_mm < t_move) {
            const double q_int = q0 + qd0 * t_mm + 0.5 * k_actual * t_mm * t_mm;
            const double g_int = (0.5 * qd0 * qd0) / (k_actual * k_actual);
            if (k_actual >= 0) {
                q_min_p = q_int; q_max_p = q_hi; gq_min_p = g_int; gq_max_p = g_hi;
            } else {
                q_min_p = q_lo; q_max_p = q_int; gq_min_p = g_lo; gq_max_p = g_int;
            }
        } else {
            q_min_p = q_lo; q_max_p = q_hi; gq_min_p = g_lo; gq_max_p = g_hi;
        }
        double v_min_p, v_max_p, gv_min_p, gv_max_p;
        if (q_dot_peak >= qd0) {
            v_min_p = qd0; v_max_p = q_dot_peak; gv_min_p = 0; gv_max_p = t_move;
        } else {
            v_min_p = q_dot_peak; v_max_p = qd0; gv_min_p = t_move; gv_max_p = 0;
        }
        double q_min_s, q_max_s, gq_min_s, gq_max_s;
        if (q_stop >= q_peak) {
            q_min_s = q_peak; q_max_s = q_stop;
            gq_min_s = 0.5 * t_move * t_move; gq_max_s = 0.5 * t_move * t_move + 0.5 * t_move * t_to_stop;
        } else {
            q_min_s = q_stop; q_max_s = q_peak;
            gq_min_s = 0.5 * t_move * t_move + 0.5 * t_move * t_to_stop; gq_max_s = 0.5 * t_move * t_move;
        }
        double v_min_s, v_max_s, gv_min_s, gv_max_s;
        if (q_dot_peak >= 0) {
            v_min_s = 0; v_max_s = q_dot_peak; gv_min_s = 0; gv_max_s = t_move;
        } else {
            v_min_s = q_dot_peak; v_max_s = 0; gv_min_s = t_move; gv_max_s = 0;
        }
        const bool a = q_min_p <= q_min_s, b = q_max_p >= q_max_s, c = v_min_p <= v_min_s, d = v_max_p >= v_max_s;
        if (ext) {
            ext[i] = a ? q_min_p : q_min_s;
            ext[i + NF] = b ? q_max_p : q_max_s;
            ext[i + 2 * NF] = c ? v_min_p : v_min_s;
            ext[i + 3 * NF] = d ? v_max_p : v_max_s;
        }
        if (grad) {  // with respect to k_actual, as the reference writes them (no k_range factor)
            grad[i] = a ? gq_min_p : gq_min_s;
            grad[i + NF] = b ? gq_max_p : gq_max_s;
            grad[i + 2 * NF] = c ? gv_min_p : gv_min_s;
            grad[i + 3 * NF] = d ? gv_max_p : gv_max_s;
        }
    }
}

void Problem::armtd_eval_g(const double* k, double* g) {
    KinematicsDynamics& K = *kd;
#pragma omp parallel for schedule(dynamic)
    for (int t = 0; t < T; t++)
        for (int l = 0; l < NJ; l++) {
            double lo[3], hi[3];
            K.links[l * T + t].slice(k, lo, hi);
            for (int e = 0; e < 3; e++) link_sliced_center[size_t(t * NJ + l) * 3 + e] = (lo[e] + hi[e]) * 0.5;
        }
    link_constraints(false, g, nullptr);
    armtd_state_extremum(k, g + NJ * T * O, nullptr);
}

void Problem::armtd_eval_jac_g(const double* k, double* values) {
    KinematicsDynamics& K = *kd;
#pragma omp parallel for schedule(dynamic)
    for (int t = 0; t < T; t++)
        for (int l = 0; l < NJ; l++) {
            double lo[3], hi[3];
            K.links[l * T + t].slice(k, lo, hi);
            for (int e = 0; e < 3; e++) link_sliced_center[size_t(t * NJ + l) * 3 + e] = (lo[e] + hi[e]) * 0.5;
            K.links[l * T + t].slice_gradient(k, &dk_link_sliced_center[size_t(t * NJ + l) * NF * 3]);
        }
    link_constraints(true, nullptr, values);
    // the reference clears only 4*NF*NF BYTES of these 4*NF rows and writes their diagonal entries (KPA/Trajectory.cu:262); on a
    // zero-filled buffer, which is what this restatement and the frozen reference outputs use, the rows are diagonal
    double* lim = values + size_t(NJ) * T * O * NF;
    std::memset(lim, 0, sizeof(double) * 4 * NF * NF);
    double gd[4 * NF];
    armtd_state_extremum(k, nullptr, gd);
    for (int r = 0; r < 4; r++)
        for (int i = 0; i < NF; i++) lim[size_t(r * NF + i) * NF + i] = gd[r * NF + i];
}

void Problem::armtd_bounds(double* g_l, double* g_u) const {
    int offset = 0;
    for (int i = 0; i < NJ * T * O; i++) {
        g_l[i] = -1e19;
        g_u[i] = 0;
    }
    offset += NJ * T * O;
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

int Problem::armtd_verdict(const double* g, int* first) const {
    auto fail = [&](int row) {
        if (first) *first = row;
        return 0;
    };
    // the reference's loop runs over NUM_FACTORS - 1 links, not NUM_JOINTS (KPA/NLPclass.cu:400): the last link is not checked
    for (int i = 0; i < NF - 1; i++)
        for (int j = 0; j < T; j++)
            for (int h = 0; h < O; h++)
                if (g[(i * T + j) * O + h] > params.collision_violation_threshold) return fail((i * T + j) * O + h);
    int offset = NJ * T * O;
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

static double wrap_pi(double angle) {  // KPA/NLPclass.cu:6-15
    double w = angle;
    while (w < -M_PI) w += 2 * M_PI;
    while (w > M_PI) w -= 2 * M_PI;
    return w;
}

double Problem::armtd_cost(const double* q_des, const double* k) const {
    double q_plan[NF];
    for (int i = 0; i < NF; i++) q_plan[i] = a_q0[i] + a_qd0[i] * 0.5 + a_k_range[i] * k[i] * 0.125;
    double obj = std::pow(wrap_pi(q_des[0] - q_plan[0]), 2) + std::pow(wrap_pi(q_des[2] - q_plan[2]), 2) +
                 std::pow(wrap_pi(q_des[4] - q_plan[4]), 2) + std::pow(wrap_pi(q_des[6] - q_plan[6]), 2) +
                 std::pow(q_des[1] - q_plan[1], 2) + std::pow(q_des[3] - q_plan[3], 2) + std::pow(q_des[5] - q_plan[5], 2);
    return obj * params.cost_scale;
}

void Problem::armtd_cost_grad(const double* q_des, const double* k, double* grad) const {
    for (int i = 0; i < NF; i++) {
        const double q_plan = a_q0[i] + a_qd0[i] * 0.5 + a_k_range[i] * k[i] * 0.125;
        const double dk = a_k_range[i] * 0.125;
        grad[i] = (i % 2 == 0) ? (2 * wrap_pi(q_plan - q_des[i]) * dk) : (2 * (q_plan - q_des[i]) * dk);
        grad[i] *= params.cost_scale;
    }
}

}  // namespace orc

using orc::Problem;
extern "C" {
int orc_armtd_build(void* h, const double* q0, const double* qd0, const double* jrs, const double* k_range, const double* obstacles,
                    int nobs, int nthreads) {
    try {
        static_cast<Problem*>(h)->build_armtd(q0, qd0, jrs, k_range, obstacles, nobs, nthreads);
    } catch (...) {
        return -1;
    }
    return 0;
}
int orc_armtd_num_constraints(void* h) { return static_cast<Problem*>(h)->armtd_num_constraints(); }
void orc_armtd_eval_g(void* h, const double* k, double* g) { static_cast<Problem*>(h)->armtd_eval_g(k, g); }
void orc_armtd_eval_jac_g(void* h, const double* k, double* v) { static_cast<Problem*>(h)->armtd_eval_jac_g(k, v); }
void orc_armtd_bounds(void* h, double* gl, double* gu) { static_cast<Problem*>(h)->armtd_bounds(gl, gu); }
int orc_armtd_verdict(void* h, const double* g, int* first) { return static_cast<Problem*>(h)->armtd_verdict(g, first); }
double orc_armtd_cost(void* h, const double* q_des, const double* k, double* grad) {
    Problem* P = static_cast<Problem*>(h);
    if (grad) P->armtd_cost_grad(q_des, k, grad);
    return P->armtd_cost(q_des, k);
}
}
