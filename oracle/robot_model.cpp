// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned: bit-identical to the reference's own sources (oracle/_ref, tests/test_oracle_pinned.py, tests/golden/ref).
// Physical constants of the Kinova Gen3 arm as used by the reference planner
// (values from KPR/KinovaWithoutGripperInfo.h:17-112 and KPR/KinovaInfo.h:17-121).
#include "robot_model.h"

#include <cstring>

namespace orc {

namespace {
const double kTrans7[8 * 3] = {0, 0,          0.15643,    //
                               0, 0.005375,   -0.12838,   //
                               0, -0.21038,   -0.006375,  //
                               0, 0.006375,   -0.21038,   //
                               0, -0.20843,   -0.006375,  //
                               0, 0.00017505, -0.10593,   //
                               0, -0.10593,   -0.00017505,
                               0, 0,          0};
const double kMass[8] = {1.3773, 1.1636, 1.1636, 0.9302, 0.6781, 0.6781, 0.5, 1.72};
const double kCom[8 * 3] = {-0.000023,  -0.010364,    -0.07336,   //
                            -0.000044,  -0.09958,     -0.013278,  //
                            -0.000044,  -0.006641,    -0.117892,  //
                            -0.000018,  -0.075478,    -0.015006,  //
                            0.000001,   -0.009432,    -0.063883,  //
                            0.000001,   -0.045483,    -0.00965,   //
                            0.000281,   0.011402,     -0.029798,  //
                            0.00000691, 0.0000044117, 0.031656};
const double kInertia[8 * 9] = {
    0.00457,   0.000001,  0.000002,  0.000001,  0.004831,  0.000448,  0.000002,  0.000448,  0.001409,  //
    0.011088,  0.000005,  0,         0.000005,  0.001072,  -0.000691, 0,         -0.000691, 0.011255,  //
    0.010932,  0,         -0.000007, 0,         0.011127,  0.000606,  -0.000007, 0.000606,  0.001043,  //
    0.008147,  -0.000001, 0,         -0.000001, 0.000631,  -0.0005,   0,         -0.0005,   0.008316,  //
    0.001596,  0,         0,         0,         0.001607,  0.000256,  0,         0.000256,  0.000399,  //
    0.001641,  0,         0,         0,         0.00041,   -0.000278, 0,         -0.000278, 0.001641,  //
    0.000587,  0.000003,  0.000003,  0.000003,  0.000369,  -0.000118, 0.000003,  -0.000118, 0.000609,  //
    0.0004596, 0,         0,         0,         0.0005181, 0,         0,         0,         0.00036051};
const double kArmature[7] = {8.03,
                             11.9962024615303644,
                             9.0025427861751517,
                             11.5806439316706360,
                             8.4665040917914123,
                             8.8537069373742430,
                             8.8587303664685315};
const double kLinkC[8][3] = {{0.000000, -0.001297, -0.088375}, {0.000000, -0.089400, -0.007877},
                             {0.000000, -0.001502, -0.129375}, {0.000000, -0.087450, -0.013648},
                             {0.000001, -0.009023, -0.071752}, {0.000000, -0.041661, -0.009251},
                             {0.000000, -0.018585, -0.033462}, {0.0, -0.00, -0.0}};
const double kLinkG[8][3] = {{0.046358, 0.047354, 0.086000}, {0.046000, 0.135400, 0.047501},
                             {0.046000, 0.047501, 0.127000}, {0.046000, 0.133450, 0.042293},
                             {0.034999, 0.044023, 0.069252}, {0.035000, 0.076739, 0.044076},
                             {0.045500, 0.056085, 0.030963}, {0.07, 0.09, 0.07}};
}  // namespace

RobotModel make_robot_model(int model_id) {
    RobotModel m;
    const bool gripper = (model_id == 1);
    m.num_joints = gripper ? 8 : 7;
    for (int i = 0; i < MAXJ; i++) m.axes[i] = (i < 7) ? 3 : 0;

    std::memset(m.trans, 0, sizeof(m.trans));
    std::memcpy(m.trans, kTrans7, sizeof(double) * 7 * 3);
    if (gripper) {
        m.trans[7 * 3 + 2] = -0.061525 - 0.10155;  // joint 8 offset; row 9 stays zero
    }
    for (int i = 0; i < MAXJ; i++) {
        const double s = (i == 0) ? 2.0 : ((i % 2 == 1) ? 1.0 : -1.0);  // pi, +pi/2, -pi/2, ...
        m.rots[i * 3 + 0] = (i == 0) ? M_PI : s * (M_PI * 0.5);
        m.rots[i * 3 + 1] = 0;
        m.rots[i * 3 + 2] = 0;
    }
    std::memcpy(m.mass, kMass, sizeof(kMass));
    std::memcpy(m.com, kCom, sizeof(kCom));
    std::memcpy(m.inertia, kInertia, sizeof(kInertia));
    for (int i = 0; i < 7; i++) m.armature[i] = kArmature[i];
    const double lb[NF] = {-1000.0, -2.41, -1000.0, -2.66, -1000.0, -2.23, -1000.0};
    const double sp[NF] = {1.3963, 1.3963, 1.3963, 1.3963, 1.2218, 1.2218, 1.2218};
    const double tq[NF] = {56.7, 56.7, 56.7, 56.7, 29.4, 29.4, 29.4};
    for (int i = 0; i < NF; i++) {
        m.state_limits_lb[i] = lb[i];
        m.state_limits_ub[i] = -lb[i];
        m.speed_limits[i] = sp[i];
        m.torque_limits[i] = tq[i];
    }
    std::memcpy(m.link_zonotope_center, kLinkC, sizeof(kLinkC));
    std::memcpy(m.link_zonotope_generators, kLinkG, sizeof(kLinkG));
    if (gripper) {
        m.alpha = 1.0;
        m.M_min = 8.29938;
        m.K = 10.0;
    }
    m.finish();
    return m;
}

}  // namespace orc
