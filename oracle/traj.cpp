// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned: bit-identical to the reference's own sources (oracle/_ref, tests/test_oracle_pinned.py, tests/golden/ref).
// See traj.h.  Line references are to the reference file KPR/Trajectory.cu.
#include "traj.h"

#include <algorithm>
#include <cmath>
#include <utility>

namespace orc {

using std::pow;

double q_des_func(double q0, double Tqd0, double TTqdd0, double k, double t) {  // :542-556
    const double B0 = -pow(t - 1, 5);
    const double B1 = 5 * t * pow(t - 1, 4);
    const double B2 = -10 * pow(t, 2) * pow(t - 1, 3);
    const double B3 = 10 * pow(t, 3) * pow(t - 1, 2);
    const double B4 = -5 * pow(t, 4) * (t - 1);
    const double B5 = pow(t, 5);
    const double beta0 = q0;
    const double beta1 = q0 + Tqd0 / 5;
    const double beta2 = q0 + (2 * Tqd0) / 5 + TTqdd0 / 20;
    const double beta3 = q0 + k;
    return B0 * beta0 + B1 * beta1 + B2 * beta2 + B3 * beta3 + B4 * beta3 + B5 * beta3;
}

double qd_des_func(double q0, double Tqd0, double TTqdd0, double k, double t) {  // :558-572
    const double dB0 = pow(t - 1.0, 4.0) * -5.0;
    const double dB1 = t * pow(t - 1.0, 3.0) * 2.0E+1 + pow(t - 1.0, 4.0) * 5.0;
    const double dB2 = t * pow(t - 1.0, 3.0) * -2.0E+1 - (t * t) * pow(t - 1.0, 2.0) * 3.0E+1;
    const double dB3 = pow(t, 3.0) * (t * 2.0 - 2.0) * 1.0E+1 + (t * t) * pow(t - 1.0, 2.0) * 3.0E+1;
    const double dB4 = pow(t, 3.0) * (t - 1.0) * -2.0E+1 - pow(t, 4.0) * 5.0;
    const double dB5 = pow(t, 4.0) * 5.0;
    const double beta0 = q0;
    const double beta1 = q0 + Tqd0 / 5;
    const double beta2 = q0 + (2 * Tqd0) / 5 + TTqdd0 / 20;
    const double beta3 = q0 + k;
    return dB0 * beta0 + dB1 * beta1 + dB2 * beta2 + dB3 * beta3 + dB4 * beta3 + dB5 * beta3;
}

double q_des_k_indep(double q0, double Tqd0, double TTqdd0, double s) {  // :812-814
    return q0 + Tqd0 * s - 6 * Tqd0 * pow(s, 3) + 8 * Tqd0 * pow(s, 4) - 3 * Tqd0 * pow(s, 5) +
           (TTqdd0 * pow(s, 2)) * 0.5 - (3 * TTqdd0 * pow(s, 3)) * 0.5 + (3 * TTqdd0 * pow(s, 4)) * 0.5 -
           (TTqdd0 * pow(s, 5)) * 0.5;
}
double qd_des_k_indep(double, double Tqd0, double TTqdd0, double s, double D) {  // :816-818
    return (pow(s - 1, 2) * (2 * Tqd0 + 4 * Tqd0 * s + 2 * TTqdd0 * s - 30 * Tqd0 * pow(s, 2) - 5 * TTqdd0 * pow(s, 2))) *
           0.5 / D;
}
double qdd_des_k_indep(double, double Tqd0, double TTqdd0, double s, double D) {  // :820-822
    return -(s - 1.0) * (TTqdd0 - (36 * Tqd0 + 8 * TTqdd0) * s + (60 * Tqd0 + 10 * TTqdd0) * pow(s, 2)) / (D * D);
}

BezierCurve::BezierCurve(const RobotModel* m, const PlannerParams* p, const double* q0_, const double* qd0_,
                         const double* qdd0_)
    : model(m), params(p) {  // :15-61
    const double D = p->duration;
    T = p->num_time_steps;
    for (int i = 0; i < NF; i++) {
        q0[i] = q0_[i];
        qd0[i] = qd0_[i];
        qdd0[i] = qdd0_[i];
        Tqd0[i] = qd0[i] * D;
        TTqdd0[i] = qdd0[i] * D * D;
    }
    const int nj = m->num_joints;
    R.resize((nj + 1) * T);
    R_t.resize(nj * T);
    qd_des.resize(NF * T);
    qda_des.resize(NF * T);
    qdda_des.resize(NF * T);
    dump.resize(NF * T);

    for (int i = 0; i < NF; i++) {
        const double a = Tqd0[i], b = TTqdd0[i];
        {
            const double sq = std::sqrt(64 * pow(a, 2) + 14 * a * b + pow(b, 2));
            q_ext_s[0][i] = (2 * a + b + sq) / (5 * (6 * a + b));
            q_ext_s[1][i] = (2 * a + b - sq) / (5 * (6 * a + b));
            q_ext_v[0][i] = q_des_k_indep(q0[i], a, b, q_ext_s[0][i]);
            q_ext_v[1][i] = q_des_k_indep(q0[i], a, b, q_ext_s[1][i]);
        }
        {
            const double sq = std::sqrt(6 * (54 * pow(a, 2) + 14 * a * b + pow(b, 2)));
            qd_ext_s[0][i] = (18 * a + 4 * b + sq) / (10 * (6 * a + b));
            qd_ext_s[1][i] = (18 * a + 4 * b - sq) / (10 * (6 * a + b));
            qd_ext_v[0][i] = qd_des_k_indep(q0[i], a, b, qd_ext_s[0][i], D);
            qd_ext_v[1][i] = qd_des_k_indep(q0[i], a, b, qd_ext_s[1][i], D);
        }
        {
            const double sq = std::sqrt(2 * (152 * pow(a, 2) + 42 * a * b + 3 * pow(b, 2)));
            qdd_ext_s[0][i] = (32 * a + 6 * b + sq) / (10 * (6 * a + b));
            qdd_ext_s[1][i] = (32 * a + 6 * b - sq) / (10 * (6 * a + b));
            qdd_ext_v[0][i] = qdd_des_k_indep(q0[i], a, b, qdd_ext_s[0][i], D);
            qdd_ext_v[1][i] = qdd_des_k_indep(q0[i], a, b, qdd_ext_s[1][i], D);
        }
    }
    ds = 1.0 / T;
}

// range of a k-independent part over [s_lb, s_ub]: endpoints plus interior extrema (:80-94 etc.)
static inline void indep_range(double v_lb, double v_ub, double s_lb, double s_ub, double e1s, double e1v, double e2s,
                               double e2v, double& radius, double& center) {
    double lb = v_lb, ub = v_ub;
    if (lb > ub) std::swap(lb, ub);
    if (s_lb < e1s && e1s < s_ub) {
        lb = std::min(lb, e1v);
        ub = std::max(ub, e1v);
    }
    if (s_lb < e2s && e2s < s_ub) {
        lb = std::min(lb, e2v);
        ub = std::max(ub, e2v);
    }
    radius = (ub - lb) * 0.5;
    center = (lb + ub) * 0.5;
}

void BezierCurve::makePolyZono(int s_ind) {  // :63-254
    const RobotModel& m = *model;
    const double D = params->duration;
    const double s_lb = s_ind * ds;
    const double s_ub = (s_ind + 1) * ds;
    const double kQddMax = 0.5 - std::sqrt(3.0) / 6;  // KPR/Trajectory.h:7-8
    const double kQddMin = 0.5 + std::sqrt(3.0) / 6;

    for (int i = 0; i < NF; i++) {
        const double kr = params->k_range[i];
        JrsDump& dmp = dump[i * T + s_ind];

        // Part 1: q_des
        double kd_lb = pow(s_lb, 3) * (6 * pow(s_lb, 2) - 15 * s_lb + 10);
        double kd_ub = pow(s_ub, 3) * (6 * pow(s_ub, 2) - 15 * s_ub + 10);
        double kd_center = (kd_ub + kd_lb) * 0.5;
        double kd_radius = (kd_ub - kd_lb) * 0.5 * kr;
        double ki_radius, q_des_center;
        indep_range(q_des_k_indep(q0[i], Tqd0[i], TTqdd0[i], s_lb), q_des_k_indep(q0[i], Tqd0[i], TTqdd0[i], s_ub),
                    s_lb, s_ub, q_ext_s[0][i], q_ext_v[0][i], q_ext_s[1][i], q_ext_v[1][i], ki_radius, q_des_center);

        const Interval qr(-kd_radius - ki_radius - m.qe, kd_radius + ki_radius + m.qe);
        const Interval kint(-kr, kr);

        // first-order Taylor expansion with interval Lagrange remainder (:103-127)
        double cos_c = std::cos(q_des_center);
        Interval cos_r = (-qr) * std::sin(q_des_center) -
                         (0.5 * cos((q_des_center + kd_center * kint) + qr)) * pow2(qr + kd_center * kint);
        cos_c += getCenter(cos_r);
        cos_r = cos_r - getCenter(cos_r);
        const double cos_coeff[2] = {-kd_center * kr * std::sin(q_des_center), getRadius(cos_r)};
        const uint64_t cos_hash[2] = {var_hash(i), var_hash(i + NF * 4)};

        double sin_c = std::sin(q_des_center);
        Interval sin_r = qr * std::cos(q_des_center) -
                         (0.5 * sin((q_des_center + kd_center * kint) + qr)) * pow2(qr + kd_center * kint);
        sin_c += getCenter(sin_r);
        sin_r = sin_r - getCenter(sin_r);
        const double sin_coeff[2] = {kd_center * kr * std::cos(q_des_center), getRadius(sin_r)};
        const uint64_t sin_hash[2] = {var_hash(i), var_hash(i + NF * 5)};

        dmp.cos_center = cos_c; dmp.cos_k = cos_coeff[0]; dmp.cos_e = cos_coeff[1];
        dmp.sin_center = sin_c; dmp.sin_k = sin_coeff[0]; dmp.sin_e = sin_coeff[1];

        PZ Ri = PZ::rpy(m.rots[i * 3], m.rots[i * 3 + 1], m.rots[i * 3 + 2]);
        if (m.axes[i] != 0) {
            Ri = Ri * PZ::rotation(cos_c, cos_coeff, cos_hash, 2, sin_c, sin_coeff, sin_hash, 2, m.axes[i]);
        }
        R[i * T + s_ind] = Ri;
        R_t[i * T + s_ind] = Ri.transpose();

        // Part 2: qd_des (:151-192)
        kd_lb = (30 * pow(s_lb, 2) * pow(s_lb - 1, 2)) / D;
        kd_ub = (30 * pow(s_ub, 2) * pow(s_ub - 1, 2)) / D;
        if (kd_ub < kd_lb) std::swap(kd_lb, kd_ub);
        kd_center = (kd_ub + kd_lb) * 0.5 * kr;
        kd_radius = (kd_ub - kd_lb) * 0.5 * kr;
        double qd_center;
        indep_range(qd_des_k_indep(q0[i], Tqd0[i], TTqdd0[i], s_lb, D), qd_des_k_indep(q0[i], Tqd0[i], TTqdd0[i], s_ub, D),
                    s_lb, s_ub, qd_ext_s[0][i], qd_ext_v[0][i], qd_ext_s[1][i], qd_ext_v[1][i], ki_radius, qd_center);
        {
            const double c1[2] = {kd_center, kd_radius + ki_radius + m.qde};
            const uint64_t h1[2] = {var_hash(i), var_hash(i + NF * 1)};
            qd_des[i * T + s_ind] = PZ::scalar_poly(qd_center, c1, h1, 2);
            const double c2[2] = {kd_center, kd_radius + ki_radius + m.qdae};
            const uint64_t h2[2] = {var_hash(i), var_hash(i + NF * 2)};
            qda_des[i * T + s_ind] = PZ::scalar_poly(qd_center, c2, h2, 2);
            dmp.qd_center = qd_center; dmp.qd_k = kd_center; dmp.qd_e = c1[1]; dmp.qda_e = c2[1];
        }

        // Part 3: qdd_des (:195-244)
        const double temp_lb = (60 * s_lb * (2 * pow(s_lb, 2) - 3 * s_lb + 1)) / D / D;
        const double temp_ub = (60 * s_ub * (2 * pow(s_ub, 2) - 3 * s_ub + 1)) / D / D;
        if (s_ub <= kQddMax) {
            kd_lb = temp_lb;
            kd_ub = temp_ub;
        } else if (s_lb <= kQddMax) {
            kd_lb = std::min(temp_lb, temp_ub);
            kd_ub = (60 * kQddMax * (2 * pow(kQddMax, 2) - 3 * kQddMax + 1)) / D / D;
        } else if (s_ub <= kQddMin) {
            kd_lb = temp_ub;
            kd_ub = temp_lb;
        } else if (s_lb <= kQddMin) {
            kd_lb = (60 * kQddMin * (2 * pow(kQddMin, 2) - 3 * kQddMin + 1)) / D / D;
            kd_ub = std::max(temp_lb, temp_ub);
        } else {
            kd_lb = temp_lb;
            kd_ub = temp_ub;
        }
        kd_center = (kd_ub + kd_lb) * 0.5 * kr;
        kd_radius = (kd_ub - kd_lb) * 0.5 * kr;
        double qdd_center;
        indep_range(qdd_des_k_indep(q0[i], Tqd0[i], TTqdd0[i], s_lb, D),
                    qdd_des_k_indep(q0[i], Tqd0[i], TTqdd0[i], s_ub, D), s_lb, s_ub, qdd_ext_s[0][i], qdd_ext_v[0][i],
                    qdd_ext_s[1][i], qdd_ext_v[1][i], ki_radius, qdd_center);
        {
            const double c3[2] = {kd_center, kd_radius + ki_radius + m.qddae};
            const uint64_t h3[2] = {var_hash(i), var_hash(i + NF * 3)};
            qdda_des[i * T + s_ind] = PZ::scalar_poly(qdd_center, c3, h3, 2);
            dmp.qdd_center = qdd_center; dmp.qdd_k = kd_center; dmp.qdd_e = c3[1];
        }
    }
    // fixed joints at the end of the chain (:247-253)
    for (int i = NF; i < m.num_joints; i++) {
        R[i * T + s_ind] = PZ::rpy(m.rots[i * 3], m.rots[i * 3 + 1], m.rots[i * 3 + 2]);
        R_t[i * T + s_ind] = R[i * T + s_ind].transpose();
    }
    R[m.num_joints * T + s_ind] = PZ::rpy(0, 0, 0);
}

// ---- closed-form extrema over the whole horizon ----------------------------------------
namespace {
struct Ext4 {
    double s[4];  // candidate locations
    double v[4];  // values there
    bool in2, in3;
};
inline Ext4 pos_candidates(double q0, double a, double b, double k) {  // :262-271
    Ext4 e;
    e.s[0] = 0;
    const double sq = std::sqrt(64 * pow(a, 2) + 14 * a * b - 120 * k * a + pow(b, 2));
    e.s[1] = (2 * a + b + sq) / (5 * (6 * a - 12 * k + b));
    e.s[2] = (2 * a + b - sq) / (5 * (6 * a - 12 * k + b));
    e.s[3] = 1;
    for (int i = 0; i < 4; i++) e.v[i] = q_des_func(q0, a, b, k, e.s[i]);
    e.in2 = (0 <= e.s[1] && e.s[1] <= 1);
    e.in3 = (0 <= e.s[2] && e.s[2] <= 1);
    return e;
}
inline Ext4 vel_candidates(double q0, double a, double b, double k) {  // :405-414
    Ext4 e;
    e.s[0] = 0;
    const double sq =
        std::sqrt(6 * (150 * pow(k, 2) - 180 * k * a - 20 * k * b + 54 * pow(a, 2) + 14 * a * b + pow(b, 2)));
    e.s[1] = (18 * a - 30 * k + 4 * b + sq) / (10 * (6 * a - 12 * k + b));
    e.s[2] = (18 * a - 30 * k + 4 * b - sq) / (10 * (6 * a - 12 * k + b));
    e.s[3] = 1;
    for (int i = 0; i < 4; i++) e.v[i] = qd_des_func(q0, a, b, k, e.s[i]);
    e.in2 = (0 <= e.s[1] && e.s[1] <= 1);
    e.in3 = (0 <= e.s[2] && e.s[2] <= 1);
    return e;
}
inline void minmax_values(const Ext4& e, double& mn, double& mx) {  // :274-283
    mn = std::min(e.v[0], e.v[3]);
    mx = std::max(e.v[0], e.v[3]);
    if (e.in2) {
        mn = std::min(mn, e.v[1]);
        mx = std::max(mx, e.v[1]);
    }
    if (e.in3) {
        mn = std::min(mn, e.v[2]);
        mx = std::max(mx, e.v[2]);
    }
}
inline void minmax_ids(const Ext4& e, int& minId, int& maxId) {  // :308-347 (ids 1..4)
    double mn, mx;
    if (e.v[0] < e.v[3]) {
        mn = e.v[0]; minId = 1; mx = e.v[3]; maxId = 4;
    } else {
        mn = e.v[3]; minId = 4; mx = e.v[0]; maxId = 1;
    }
    if (e.in2) {
        if (e.v[1] < mn) { mn = e.v[1]; minId = 2; }
        if (mx < e.v[1]) { mx = e.v[1]; maxId = 2; }
    }
    if (e.in3) {
        if (e.v[2] < mn) { mn = e.v[2]; minId = 3; }
        if (mx < e.v[2]) { mx = e.v[2]; maxId = 3; }
    }
}
// d/dk of q_des evaluated at an interior stationary point s*(k).  The reference evaluates the
// symbolic total derivative (generated code, :601-685); by the envelope theorem the dq/ds * ds*/dk
// term vanishes at a stationary point, leaving dq/dk = B3+B4+B5 = s^3 (6 s^2 - 15 s + 10).
inline double dpos_dk_at(double s) { return 10 * pow(s, 3) * pow(s - 1, 2) - 5 * pow(s, 4) * (s - 1) + pow(s, 5); }
// same for qd_des (:687-810): d(qd)/dk = dB3+dB4+dB5 = 30 s^2 (s-1)^2
inline double dvel_dk_at(double s) {
    return (pow(s, 3.0) * (s * 2.0 - 2.0) * 1.0E+1 + (s * s) * pow(s - 1.0, 2.0) * 3.0E+1) +
           (pow(s, 3.0) * (s - 1.0) * -2.0E+1 - pow(s, 4.0) * 5.0) + pow(s, 4.0) * 5.0;
}
}  // namespace

void BezierCurve::jointPositionExtremum(double* ext, const double* k) const {
    for (int i = 0; i < NF; i++) {
        const double ka = params->k_range[i] * k[i];
        const Ext4 e = pos_candidates(q0[i], Tqd0[i], TTqdd0[i], ka);
        minmax_values(e, ext[i], ext[i + NF]);
    }
}
void BezierCurve::jointVelocityExtremum(double* ext, const double* k) const {
    for (int i = 0; i < NF; i++) {
        const double ka = params->k_range[i] * k[i];
        const Ext4 e = vel_candidates(q0[i], Tqd0[i], TTqdd0[i], ka);
        double mn, mx;
        minmax_values(e, mn, mx);
        ext[i] = mn / params->duration;
        ext[i + NF] = mx / params->duration;
    }
}
void BezierCurve::jointPositionExtremumGradient(double* g, const double* k) const {
    for (int i = 0; i < NF; i++) {
        const double ka = params->k_range[i] * k[i];
        const Ext4 e = pos_candidates(q0[i], Tqd0[i], TTqdd0[i], ka);
        int mnId, mxId;
        minmax_ids(e, mnId, mxId);
        auto grad = [&](int id) { return id == 1 ? 0.0 : id == 4 ? 1.0 : dpos_dk_at(e.s[id - 1]); };
        for (int j = 0; j < NF; j++) {
            g[i * NF + j] = (i == j) ? grad(mnId) * params->k_range[i] : 0.0;
            g[(i + NF) * NF + j] = (i == j) ? grad(mxId) * params->k_range[i] : 0.0;
        }
    }
}
void BezierCurve::jointVelocityExtremumGradient(double* g, const double* k) const {
    for (int i = 0; i < NF; i++) {
        const double ka = params->k_range[i] * k[i];
        const Ext4 e = vel_candidates(q0[i], Tqd0[i], TTqdd0[i], ka);
        int mnId, mxId;
        minmax_ids(e, mnId, mxId);
        // sic: the reference returns 1.0 for the s = 1 candidate (:505-507) although qd(1) = 0 for every k
        auto grad = [&](int id) { return id == 1 ? 0.0 : id == 4 ? 1.0 : dvel_dk_at(e.s[id - 1]); };
        for (int j = 0; j < NF; j++) {
            g[i * NF + j] = (i == j) ? grad(mnId) * params->k_range[i] / params->duration : 0.0;
            g[(i + NF) * NF + j] = (i == j) ? grad(mxId) * params->k_range[i] / params->duration : 0.0;
        }
    }
}

}  // namespace orc
