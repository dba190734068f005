// ORACLE — TEST INFRASTRUCTURE ONLY.  Constants checked through every pinned quantity (see planner.h).
//
// Robot and planner constants of the ARMOUR Kinova Gen3 planner, restated as a
// runtime struct so that the same oracle binary can serve the 7-joint model
// (reference KPR/KinovaWithoutGripperInfo.h:10-112) and the 8-joint model with
// a fixed gripper link (reference KPR/KinovaInfo.h:10-121), and so that the
// planner knobs of KPR/Parameters.h:10-58 (threshold, k_range, time steps) can
// be varied by tests.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use anything under oracle/.
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

constexpr int NF = 7;        // NUM_FACTORS: trajectory parameters / actuated joints
constexpr int MAXJ = 8;      // largest NUM_JOINTS we support (7 arm links + fixed gripper)
constexpr int NVAR = NF * 6; // k, qde, qdae, qddae, cosqe, sinqe  (KPR/PZsparse.h:6-20)

struct RobotModel {
    int num_joints = 7;
    int axes[MAXJ] = {3, 3, 3, 3, 3, 3, 3, 0};
    double trans[(MAXJ + 1) * 3] = {0};
    double rots[MAXJ * 3] = {0};
    double mass[MAXJ] = {0};
    double mass_uncertainty = 0.03;
    double com[MAXJ * 3] = {0};
    double com_uncertainty = 0.0;
    double inertia[MAXJ * 9] = {0};
    double inertia_uncertainty = 0.03;
    double friction[MAXJ] = {0};
    double damping[MAXJ] = {0};
    double armature[MAXJ] = {0};
    double state_limits_lb[NF] = {0};
    double state_limits_ub[NF] = {0};
    double speed_limits[NF] = {0};
    double torque_limits[NF] = {0};
    double gravity = 9.81;
    double link_zonotope_center[MAXJ][3] = {{0}};
    double link_zonotope_generators[MAXJ][3] = {{0}};
    // ultimate-bound constants
    double alpha = 10.0, V_m = 1e-2, M_max = 15.79635774, M_min = 5.095620491878957, K = 5.0;
    double eps = 0, qe = 0, qde = 0, qdae = 0, qddae = 0;

    void finish() {  // KPR/KinovaWithoutGripperInfo.h:107-111
        eps = std::sqrt(2 * V_m / M_min);
        qe = eps / K;
        qde = 2 * eps;
        qdae = eps;
        qddae = 2 * K * eps;
    }
};

struct PlannerParams {  // KPR/Parameters.h
    double simplify_threshold = 5e-4;
    double duration = 1.0;
    int num_time_steps = 128;
    double k_range[NF] = {M_PI / 48, M_PI / 48, M_PI / 48, M_PI / 48, M_PI / 48, M_PI / 48, M_PI / 48};
    double t_plan = 0.5;  // KPR/armour_main.cu:80
    double collision_violation_threshold = 1e-4;
    double torque_violation_threshold = 1e-2;
    double cost_scale = 10.0;
    int max_obstacles = 40;
};

// model_id 0: Kinova Gen3 without gripper (the shipped configuration);
// model_id 1: Kinova Gen3 with the fixed 1.72 kg gripper link.
RobotModel make_robot_model(int model_id);

}  // namespace orc
