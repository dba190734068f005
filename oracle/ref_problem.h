// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Orchestration of the reference's main() up to the reach sets, shared by the two drivers around the REFERENCE's
// own classes (ref_driver.cpp: host-only build with g++; ref_driver_nlp.cu: full build with nvcc, including
// KPR/CollisionChecking.cu and KPR/NLPclass.cu).  Restates KPR/armour_main.cu:96-142 (JRS, FK, reduce_link_PZ,
// nominal + interval RNEA, disturbance, reduce) and :172-201 (robust-input radius); every arithmetic statement
// executed is the reference's own code.
#pragma once
#include <cstring>
#include <vector>

#include "Dynamics.h"  // the reference's header (-I <reference>/kinova_planner_realtime)

namespace {
struct RefProblem {
    BezierCurve traj;
    KinematicsDynamics kd;
    std::vector<Eigen::MatrixXd> link_gens;  // [t*NUM_JOINTS + l], 3x6
    Eigen::MatrixXd torque_radius;           // (NUM_FACTORS, NUM_TIME_STEPS)
};

// returns false if the reference threw
inline bool ref_problem_build(RefProblem* P, const double* q0, const double* qd0, const double* qdd0, int nthreads) {
    Eigen::VectorXd a(NUM_FACTORS), b(NUM_FACTORS), c(NUM_FACTORS);
    for (int i = 0; i < NUM_FACTORS; i++) {
        a(i) = q0[i];
        b(i) = qd0[i];
        c(i) = qdd0[i];
    }
    if (nthreads > 0) omp_set_num_threads(nthreads);
    try {
        P->traj = BezierCurve(a, b, c);
        int t = 0;
#pragma omp parallel for shared(P) private(t) schedule(dynamic, 1)
        for (t = 0; t < NUM_TIME_STEPS; t++) P->traj.makePolyZono(t);  // armour_main.cu:99-102

        P->kd = KinematicsDynamics(&P->traj);
        P->link_gens.resize(NUM_TIME_STEPS * NUM_JOINTS);
#pragma omp parallel for shared(P) private(t) schedule(dynamic)
        for (t = 0; t < NUM_TIME_STEPS; t++) {  // armour_main.cu:117-142
            KinematicsDynamics& kd = P->kd;
            kd.fk(t);
            for (int i = 0; i < NUM_JOINTS; i++) P->link_gens[t * NUM_JOINTS + i] = kd.links(i, t).reduce_link_PZ();
            kd.rnea_nominal(t);
            kd.rnea_interval(t);
            for (int i = 0; i < NUM_FACTORS; i++) kd.u_nom_int(i, t) = kd.u_nom_int(i, t) - kd.u_nom(i, t);
            for (int i = 0; i < NUM_FACTORS; i++) kd.u_nom(i, t).reduce();
        }

        P->torque_radius = Eigen::MatrixXd::Zero(NUM_FACTORS, NUM_TIME_STEPS);
        for (int t_ind = 0; t_ind < NUM_TIME_STEPS; t_ind++) {  // armour_main.cu:177-201
            Interval rho = Interval(0.0);
            for (int i = 0; i < NUM_FACTORS; i++) {
                MatrixXInt w = P->kd.u_nom_int(i, t_ind).toInterval();
                rho += w(0) * w(0);
                P->torque_radius(i, t_ind) =
                    alpha * (M_max - M_min) * eps + 0.5 * std::max(std::abs(w(0).lower()), std::abs(w(0).upper()));
            }
            rho = sqrt(rho);
            for (int i = 0; i < NUM_FACTORS; i++) P->torque_radius(i, t_ind) += 0.5 * rho.upper();
            for (int i = 0; i < NUM_FACTORS; i++) P->torque_radius(i, t_ind) += P->kd.u_nom(i, t_ind).independent(0);
            for (int i = 0; i < NUM_FACTORS; i++) P->torque_radius(i, t_ind) += friction[i];
        }
    } catch (...) {
        return false;
    }
    P->kd.traj = &P->traj;
    return true;
}
}  // namespace
