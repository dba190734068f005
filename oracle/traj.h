// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned: bit-identical to the reference's own sources (oracle/_ref, tests/test_oracle_pinned.py, tests/golden/ref).
//
// Degree-5 Bezier desired trajectory and its per-interval joint reachable set (JRS),
// restating KPR/Trajectory.h:10-95 and KPR/Trajectory.cu:15-822.
#pragma once
#include <vector>

#include "pz.h"

namespace orc {

struct JrsDump {  // raw numbers of one (joint, interval), for kernel-level parity tests
    double cos_center, cos_k, cos_e;   // cos(q_des) = cos_center + cos_k*k_i + cos_e*cosqe_i   (before simplify)
    double sin_center, sin_k, sin_e;
    double qd_center, qd_k, qd_e, qda_e;
    double qdd_center, qdd_k, qdd_e;
};

struct BezierCurve {
    const RobotModel* model = nullptr;
    const PlannerParams* params = nullptr;
    double q0[NF], qd0[NF], qdd0[NF], Tqd0[NF], TTqdd0[NF];
    double q_ext_s[2][NF], q_ext_v[2][NF];      // interior extrema (location, value) of the k-independent part
    double qd_ext_s[2][NF], qd_ext_v[2][NF];
    double qdd_ext_s[2][NF], qdd_ext_v[2][NF];
    double ds = 0;
    int T = 0;

    // per (joint, t): index i*T + t ; R has num_joints+1 rows
    std::vector<PZ> R, R_t, qd_des, qda_des, qdda_des;
    std::vector<JrsDump> dump;

    BezierCurve(const RobotModel* m, const PlannerParams* p, const double* q0, const double* qd0, const double* qdd0);
    void makePolyZono(int s_ind);  // KPR/Trajectory.cu:63-254

    void jointPositionExtremum(double* ext /*[2*NF]*/, const double* k) const;          // :256-288
    void jointPositionExtremumGradient(double* grad /*[2*NF*NF]*/, const double* k) const;  // :290-397
    void jointVelocityExtremum(double* ext, const double* k) const;                      // :399-431
    void jointVelocityExtremumGradient(double* grad, const double* k) const;             // :433-540
};

double q_des_func(double q0, double Tqd0, double TTqdd0, double k, double t);    // :542-556
double qd_des_func(double q0, double Tqd0, double TTqdd0, double k, double t);   // :558-572
double q_des_k_indep(double q0, double Tqd0, double TTqdd0, double s);           // :812-814
double qd_des_k_indep(double q0, double Tqd0, double TTqdd0, double s, double duration);   // :816-818
double qdd_des_k_indep(double q0, double Tqd0, double TTqdd0, double s, double duration);  // :820-822

}  // namespace orc
