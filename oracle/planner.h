// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned against the reference's own sources: reach-set build, slices and Bezier rows through oracle/_ref/libarmour_ref.so (tests/test_oracle_pinned.py); collision rows, bounds, cost and verdict through the reference's CUDA kernels and armtd_NLP compiled by nvcc (oracle/_ref/libarmour_ref_cuda.so, run on a B200, frozen as tests/golden/refcuda/, tests/test_refcuda_golden.py).
//
// One planning problem end to end on the CPU: reach-set build (KPR/armour_main.cu:86-216),
// collision hyper-planes (KPR/CollisionChecking.cu:26-39,136-228), the NLP rows and Jacobian
// (KPR/NLPclass.cu:45-396, KPR/CollisionChecking.cu:230-299), bounds and the feasibility verdict
// (KPR/NLPclass.cu:116-165,449-537), cost and cost gradient (:207-268).
#pragma once
#include <memory>
#include <vector>

#include "dynamics.h"

namespace orc {

constexpr int kBufGen = 9;   // 3 obstacle generators + 3 link generators + 3 radius generators
constexpr int kComb = 36;    // C(9,2)

struct Problem {
    RobotModel model;
    PlannerParams params;
    int T = 0, NJ = 0, O = 0;
    std::unique_ptr<BezierCurve> traj;
    std::unique_ptr<KinematicsDynamics> kd;
    std::vector<double> link_gens;      // [T*NJ][18] column-major 3x6        (armour_main.cu:113,126)
    std::vector<double> torque_radius;  // [NF*T], index j*T + t               (armour_main.cu:168-201)
    std::vector<double> obstacles;      // [O*12]: c, g1, g2, g3
    std::vector<double> A, d, delta;    // [((t*NJ + l)*O + o)*36 + p] (A has 3 per entry)
    std::vector<double> link_sliced_center;     // [T*NJ*3]   (NLPclass.h:150)
    std::vector<double> dk_link_sliced_center;  // [T*NJ*NF*3]
    Stats stats;
    double build_ms = 0;

    Problem(int model_id, const PlannerParams& p);
    // armour_main.cu sections II.A-II.D.  nthreads <= 0: OpenMP default.
    void build(const double* q0, const double* qd0, const double* qdd0, const double* obstacles, int nobs, int nthreads);
    int num_constraints() const { return NF * T + NJ * T * O + NF * 4; }
    void eval_g(const double* k, double* g);                 // NLPclass.cu:272-324
    void eval_jac_g(const double* k, double* values);        // NLPclass.cu:330-396
    void bounds(double* g_l, double* g_u) const;             // NLPclass.cu:116-165
    // verdict of finalize_solution: returns 1 feasible / 0 infeasible; first violated row or -1
    int verdict(const double* g, int* first_violation) const;  // NLPclass.cu:449-537
    double cost(const double* q_des, const double* k) const;    // NLPclass.cu:207-236
    void cost_grad(const double* q_des, const double* k, double* grad) const;  // :241-268

    // ---- ARMTD comparison planner (kinova_planner_realtime_armtd_comparison, "KPA"; oracle/armtd.cpp): constant-acceleration
    // trajectory with an offline joint reachable set, forward kinematics only, no torque rows
    double a_q0[NF] = {0}, a_qd0[NF] = {0}, a_k_range[NF] = {0};
    // jrs: [6][NF][T] = c_cos, g_cos, r_cos, c_sin, g_sin, r_sin (the input file's arrays, KPA/armtd_main.cu:70-88)
    void build_armtd(const double* q0, const double* qd0, const double* jrs, const double* k_range, const double* obstacles,
                     int nobs, int nthreads);                                                  // armtd_main.cu:107-160
    int armtd_num_constraints() const { return NJ * T * O + NF * 4; }                         // KPA/NLPclass.cu:43-44
    void armtd_eval_g(const double* k, double* g);                                             // :248-283
    void armtd_eval_jac_g(const double* k, double* values);                                    // :288-336
    void armtd_bounds(double* g_l, double* g_u) const;                                         // :75-142
    int armtd_verdict(const double* g, int* first_violation) const;                            // :366-455
    double armtd_cost(const double* q_des, const double* k) const;                             // :178-212
    void armtd_cost_grad(const double* q_des, const double* k, double* grad) const;            // :217-243
    // ConstantAccelerationCurve::returnJointStateExtremum[Gradient] (KPA/Trajectory.cu:88-384): ext[4*NF], grad[4*NF] (diagonal)
    void armtd_state_extremum(const double* k, double* ext, double* grad_diag) const;

private:
    void init_hyperplanes();
    void link_constraints(bool with_grad, double* link_c, double* grad_link_c);
};

}  // namespace orc
