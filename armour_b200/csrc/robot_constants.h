// Robot / planner constants for the B200 hot path.
//
// Values are the physical parameters of the Kinova Gen3 arm used by the reference planner
// (reference KPR/KinovaWithoutGripperInfo.h:10-112 for the 7-joint model, KPR/KinovaInfo.h:10-121
// for the model with the fixed gripper link) and its planner knobs (KPR/Parameters.h:10-58).
// They are carried in one POD struct that lives in __constant__ memory on the device, so the
// reference's compile-time #defines become run-time configuration (threshold, k_range, number of
// time steps, obstacle capacity, uncertainty percentages).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace armour {

constexpr int NF = 7;    // trajectory parameters / actuated joints (reference NUM_FACTORS)
constexpr int MAXJ = 8;  // largest supported NUM_JOINTS
constexpr int NCOMB = 36;  // C(9,2) generator pairs of a buffered obstacle zonotope

struct RobotConstants {
    int num_joints;
    int num_time_steps;
    int axes[MAXJ];
    double trans[(MAXJ + 1) * 3];
    double rots[MAXJ * 3];
    double rrpy[MAXJ * 9];  // constant joint-frame rotations Rrpy(rots), column-major (KPR/PZsparse.cu:160-176)
    double mass[MAXJ];
    double com[MAXJ * 3];
    double inertia[MAXJ * 9];
    double friction[MAXJ], damping[MAXJ], armature[MAXJ];
    double state_limits_lb[NF], state_limits_ub[NF], speed_limits[NF], torque_limits[NF];
    double link_zonotope_center[MAXJ * 3], link_zonotope_generators[MAXJ * 3];
    double gravity;
    double mass_uncertainty, com_uncertainty, inertia_uncertainty;
    double alpha, V_m, M_max, M_min, K, eps, qe, qde, qdae, qddae;
    // planner parameters
    double simplify_threshold;
    double duration;
    double k_range[NF];
    double t_plan;
    double collision_violation_threshold, torque_violation_threshold, cost_scale;
};

inline RobotConstants make_robot_constants(int model_id) {
    RobotConstants m;
    std::memset(&m, 0, sizeof(m));
    const bool gripper = (model_id == 1);
    m.num_joints = gripper ? 8 : 7;
    m.num_time_steps = 128;
    for (int i = 0; i < MAXJ; i++) m.axes[i] = (i < 7) ? 3 : 0;
    const double trans7[7 * 3] = {0, 0,          0.15643,   0, 0.005375, -0.12838,  0, -0.21038,   -0.006375,
                                  0, 0.006375,   -0.21038,  0, -0.20843, -0.006375, 0, 0.00017505, -0.10593,
                                  0, -0.10593,   -0.00017505};
    std::memcpy(m.trans, trans7, sizeof(trans7));
    if (gripper) m.trans[7 * 3 + 2] = -0.061525 - 0.10155;
    m.rots[0] = M_PI;
    for (int i = 1; i < MAXJ; i++) m.rots[i * 3] = (i % 2 == 1) ? M_PI * 0.5 : -M_PI * 0.5;
    for (int i = 0; i < MAXJ; i++) {
        const double roll = m.rots[i * 3], pitch = m.rots[i * 3 + 1], yaw = m.rots[i * 3 + 2];
        double* R = &m.rrpy[i * 9];
        using std::cos;
        using std::sin;
        R[0 + 0 * 3] = cos(pitch) * cos(yaw);
        R[0 + 1 * 3] = -cos(pitch) * sin(yaw);
        R[0 + 2 * 3] = sin(pitch);
        R[1 + 0 * 3] = cos(roll) * sin(yaw) + cos(yaw) * sin(pitch) * sin(roll);
        R[1 + 1 * 3] = cos(roll) * cos(yaw) - sin(pitch) * sin(roll) * sin(yaw);
        R[1 + 2 * 3] = -cos(pitch) * sin(roll);
        R[2 + 0 * 3] = sin(roll) * sin(yaw) - cos(roll) * cos(yaw) * sin(pitch);
        R[2 + 1 * 3] = cos(yaw) * sin(roll) + cos(roll) * sin(pitch) * sin(yaw);
        R[2 + 2 * 3] = cos(pitch) * cos(roll);
    }
    const double mass[8] = {1.3773, 1.1636, 1.1636, 0.9302, 0.6781, 0.6781, 0.5, 1.72};
    std::memcpy(m.mass, mass, sizeof(mass));
    const double com[8 * 3] = {-0.000023, -0.010364, -0.07336,  -0.000044,  -0.09958,     -0.013278,
                               -0.000044, -0.006641, -0.117892, -0.000018,  -0.075478,    -0.015006,
                               0.000001,  -0.009432, -0.063883, 0.000001,   -0.045483,    -0.00965,
                               0.000281,  0.011402,  -0.029798, 0.00000691, 0.0000044117, 0.031656};
    std::memcpy(m.com, com, sizeof(com));
    const double inertia[8 * 9] = {
        0.00457,   0.000001,  0.000002,  0.000001,  0.004831,  0.000448,  0.000002,  0.000448,  0.001409,
        0.011088,  0.000005,  0,         0.000005,  0.001072,  -0.000691, 0,         -0.000691, 0.011255,
        0.010932,  0,         -0.000007, 0,         0.011127,  0.000606,  -0.000007, 0.000606,  0.001043,
        0.008147,  -0.000001, 0,         -0.000001, 0.000631,  -0.0005,   0,         -0.0005,   0.008316,
        0.001596,  0,         0,         0,         0.001607,  0.000256,  0,         0.000256,  0.000399,
        0.001641,  0,         0,         0,         0.00041,   -0.000278, 0,         -0.000278, 0.001641,
        0.000587,  0.000003,  0.000003,  0.000003,  0.000369,  -0.000118, 0.000003,  -0.000118, 0.000609,
        0.0004596, 0,         0,         0,         0.0005181, 0,         0,         0,         0.00036051};
    std::memcpy(m.inertia, inertia, sizeof(inertia));
    const double armature[7] = {8.03,
                                11.9962024615303644,
                                9.0025427861751517,
                                11.5806439316706360,
                                8.4665040917914123,
                                8.8537069373742430,
                                8.8587303664685315};
    std::memcpy(m.armature, armature, sizeof(armature));
    const double lb[NF] = {-1000.0, -2.41, -1000.0, -2.66, -1000.0, -2.23, -1000.0};
    const double sp[NF] = {1.3963, 1.3963, 1.3963, 1.3963, 1.2218, 1.2218, 1.2218};
    const double tq[NF] = {56.7, 56.7, 56.7, 56.7, 29.4, 29.4, 29.4};
    for (int i = 0; i < NF; i++) {
        m.state_limits_lb[i] = lb[i];
        m.state_limits_ub[i] = -lb[i];
        m.speed_limits[i] = sp[i];
        m.torque_limits[i] = tq[i];
    }
    const double lc[8 * 3] = {0.000000, -0.001297, -0.088375, 0.000000, -0.089400, -0.007877, 0.000000, -0.001502,
                              -0.129375, 0.000000, -0.087450, -0.013648, 0.000001, -0.009023, -0.071752, 0.000000,
                              -0.041661, -0.009251, 0.000000, -0.018585, -0.033462, 0.0,      -0.00,     -0.0};
    const double lg[8 * 3] = {0.046358, 0.047354, 0.086000, 0.046000, 0.135400, 0.047501, 0.046000, 0.047501,
                              0.127000, 0.046000, 0.133450, 0.042293, 0.034999, 0.044023, 0.069252, 0.035000,
                              0.076739, 0.044076, 0.045500, 0.056085, 0.030963, 0.07,     0.09,     0.07};
    std::memcpy(m.link_zonotope_center, lc, sizeof(lc));
    std::memcpy(m.link_zonotope_generators, lg, sizeof(lg));
    m.gravity = 9.81;
    m.mass_uncertainty = 0.03;
    m.com_uncertainty = 0.0;
    m.inertia_uncertainty = 0.03;
    m.alpha = gripper ? 1.0 : 10.0;
    m.V_m = 1e-2;
    m.M_max = 15.79635774;
    m.M_min = gripper ? 8.29938 : 5.095620491878957;
    m.K = gripper ? 10.0 : 5.0;
    m.eps = std::sqrt(2 * m.V_m / m.M_min);
    m.qe = m.eps / m.K;
    m.qde = 2 * m.eps;
    m.qdae = m.eps;
    m.qddae = 2 * m.K * m.eps;
    m.simplify_threshold = 5e-4;
    m.duration = 1.0;
    for (int i = 0; i < NF; i++) m.k_range[i] = M_PI / 48;
    m.t_plan = 0.5;
    m.collision_violation_threshold = 1e-4;
    m.torque_violation_threshold = 1e-2;
    m.cost_scale = 10.0;
    return m;
}

}  // namespace armour
