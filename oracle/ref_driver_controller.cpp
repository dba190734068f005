// ORACLE — TEST INFRASTRUCTURE ONLY.
// C entry points around the REFERENCE's own robust-controller sources (MEX/spatial.cpp, spatial_interval.cpp,
// robot_models.cpp, rnea.cpp, robust_controller.cpp, compiled where they lie by oracle/Makefile.ref target `mex` against
// the stand-in Eigen / Boost.Interval headers): what MEX/kinova_controller.cpp does between its mxGetData and
// mxCreateNumericMatrix calls, without MATLAB.  One translation unit, like each of the reference's own (its headers
// include the .cpp files of the spatial classes).
#include "robust_controller.cpp"
#include "rnea.cpp"
#include "robot_models.cpp"

namespace {
struct RefController {
    Robot* robot;
};
Eigen::VectorXd vec(const double* p, int n) {
    Eigen::VectorXd v(n);
    for (int i = 0; i < n; i++) v(i) = p[i];
    return v;
}
}  // namespace

extern "C" {

void* refctl_create(const char* model_file, double eps) {
    try {
        RefController* h = new RefController;
        h->robot = new Robot(std::string(model_file), eps);  // MEX/kinova_controller.cpp:29
        return h;
    } catch (...) {
        return nullptr;
    }
}
void refctl_destroy(void* h) {
    if (!h) return;
    delete static_cast<RefController*>(h)->robot;
    delete static_cast<RefController*>(h);
}
int refctl_num_joints(void* h) { return static_cast<RefController*>(h)->robot->numJoints; }

// passRNEA (MEX/rnea.cpp:6-94)
void refctl_rnea(void* h, const double* q, const double* qd, const double* qda, const double* qdd, int friction, int gravity,
                 double* tau) {
    Robot* R = static_cast<RefController*>(h)->robot;
    const int n = R->numJoints;
    Eigen::VectorXd vq = vec(q, n), vqd = vec(qd, n), vqda = vec(qda, n), vqdd = vec(qdd, n), t(n);
    passRNEA(t, R->RobotModelPtr, vq, vqd, vqda, vqdd, friction != 0, gravity != 0);
    for (int i = 0; i < n; i++) tau[i] = t(i);
}
// passRNEA_Int (MEX/rnea.cpp:96-187)
void refctl_rnea_int(void* h, const double* q, const double* qd, const double* qda, const double* qdd, int friction,
                     int gravity, double* lo, double* hi) {
    Robot* R = static_cast<RefController*>(h)->robot;
    const int n = R->numJoints;
    Eigen::VectorXd vq = vec(q, n), vqd = vec(qd, n), vqda = vec(qda, n), vqdd = vec(qdd, n);
    VectorXint t(n);
    passRNEA_Int(t, R->IntRobotModelPtr, vq, vqd, vqda, vqdd, friction != 0, gravity != 0);
    for (int i = 0; i < n; i++) {
        lo[i] = t(i).lower();
        hi[i] = t(i).upper();
    }
}
// RobustController::update, ARMOUR method, as MEX/kinova_controller.cpp:36-83 sets it up (Kr diagonal, no friction).
// Returns 0, or 1 if the reference threw ("Nominal model output falls outside interval output").
int refctl_update(void* h, const double* Kr_diag, double alpha, double V_max, double r_norm_threshold, const double* q,
                  const double* qd, const double* q_des, const double* qd_des, const double* qdd_des, double* u,
                  double* u_nominal, double* v) {
    Robot* R = static_cast<RefController*>(h)->robot;
    const int n = R->numJoints;
    Eigen::MatrixXd Kr = Eigen::MatrixXd::Identity(n, n);
    for (int i = 0; i < n; i++) Kr(i, i) = Kr_diag[i];
    RobustController c(Kr, alpha, V_max, r_norm_threshold);
    c.applyFriction = false;
    Eigen::VectorXd vq = vec(q, n), vqd = vec(qd, n), vqdes = vec(q_des, n), vqddes = vec(qd_des, n), vqdddes = vec(qdd_des, n);
    try {
        Eigen::VectorXd out = c.update(R, vq, vqd, vqdes, vqddes, vqdddes);
        for (int i = 0; i < n; i++) {
            u[i] = out(i);
            u_nominal[i] = c.u_nominal(i);
            v[i] = c.v(i);
        }
    } catch (...) {
        return 1;
    }
    return 0;
}
// the interval model after the conversion of the constructor (MEX/robot_models.cpp:129-156, 175-237), for the oracle's pin:
// per joint S.w[3] S.v[3] | XTree.R[9] (row major) XTree.p[3] | m | I_bar[9] | m_c_hat[9], each as (lower, upper)
void refctl_int_model(void* h, double* out) {
    IntModel* M = static_cast<RefController*>(h)->robot->IntRobotModelPtr;
    int k = 0;
    auto put = [&](const Interval& x) {
        out[k++] = x.lower();
        out[k++] = x.upper();
    };
    for (int i = 0; i < M->numJoints; i++) {
        for (int a = 0; a < 3; a++) put(M->S[i].w(a));
        for (int a = 0; a < 3; a++) put(M->S[i].v(a));
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) put(M->XTree[i].R(a, b));
        for (int a = 0; a < 3; a++) put(M->XTree[i].p(a));
        put(M->I[i].m);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) put(M->I[i].I_bar(a, b));
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) put(M->I[i].m_c_hat(a, b));
    }
}
}
