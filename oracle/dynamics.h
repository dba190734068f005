// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned: bit-identical to the reference's own sources (oracle/_ref, tests/test_oracle_pinned.py, tests/golden/ref).
//
// PZ forward kinematics and PZ recursive Newton-Euler, restating KPR/Dynamics.h:6-48 and
// KPR/Dynamics.cu:6-181.
#pragma once
#include <functional>
#include <vector>

#include "traj.h"

namespace orc {

struct KinematicsDynamics {
    const RobotModel* model = nullptr;
    BezierCurve* traj = nullptr;
    int T = 0;
    std::vector<PZ> mass_nominal, mass_uncertain, I_nominal, I_uncertain;
    std::vector<PZ> links;      // [num_joints * T]
    std::vector<PZ> u_nom;      // [NF * T]
    std::vector<PZ> u_nom_int;  // [NF * T]
    // optional debugging hook: called with (name, joint, value) for the named intermediates of fk / rnea
    std::function<void(const char*, int, const PZ&)> probe;

    explicit KinematicsDynamics(BezierCurve* traj);
    void fk(int t);                                                           // :69-81
    void rnea(int t, const std::vector<PZ>& mass_arr, const std::vector<PZ>& I_arr, std::vector<PZ>& u);  // :83-181
    void rnea_nominal(int t) { rnea(t, mass_nominal, I_nominal, u_nom); }
    void rnea_interval(int t) { rnea(t, mass_uncertain, I_uncertain, u_nom_int); }
};

}  // namespace orc
