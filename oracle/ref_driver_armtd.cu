// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// C driver around the REFERENCE's ARMTD comparison planner: its own PZsparse.cu, Trajectory.cu
// (ConstantAccelerationCurve), Dynamics.cu (forward kinematics only), CollisionChecking.cu and NLPclass.cu under
// kinova_planner_realtime_armtd_comparison/ ("KPA") are compiled by nvcc from where they lie (oracle/Makefile.ref, target
// `armtd` -> _ref/libarmour_ref_armtd.so) against the stand-in headers of oracle/ref_shim/.  The driver restates only what
// KPA/armtd_main.cu does around those classes (:107-160: trajectory, Obstacles, KinematicsDynamics, makePolyZono, fk,
// reduce_link_PZ, initializeHyperPlane; :171-172: armtd_NLP::set_parameters) and calls the armtd_NLP members the way Ipopt
// would.  Needs a GPU to run (the reference's collision kernels): tools/make_golden_armtd.py freezes its outputs as
// tests/golden/armtd/reference.npz.
#include "NLPclass.h"  // KPA's header: Dynamics.h, CollisionChecking.h, armtd_NLP

#include <cstring>
#include <vector>

namespace {
struct RefArmtd {
    std::vector<double> q0, qd0, q_des, jrs[6], k_range;  // storage the reference's classes point into
    double obstacles[MAX_OBSTACLE_NUM * (MAX_OBSTACLE_GENERATOR_NUM + 1) * 3];
    int nobs = 0;
    ConstantAccelerationCurve traj;
    Obstacles* O = nullptr;
    KinematicsDynamics* kd = nullptr;
    armtd_NLP* nlp = nullptr;
    Eigen::Matrix<double, 3, 3 + 3>* gens = nullptr;
    ~RefArmtd() {
        delete nlp;
        delete kd;
        delete O;
        delete[] gens;
    }
};
}  // namespace

extern "C" {

int refarmtd_num_time_steps() { return NUM_TIME_STEPS; }
int refarmtd_num_joints() { return NUM_JOINTS; }

// q0, qd0, q_des, k_range [7]; jrs: the six arrays of the input file in its order (c_cos, g_cos, r_cos, c_sin, g_sin, r_sin),
// each [7][NUM_TIME_STEPS] joint-major (KPA/armtd_main.cu:70-88); obstacles [nobs*12]
void* refarmtd_build(const double* q0, const double* qd0, const double* q_des, const double* jrs, const double* k_range,
                     const double* obstacles, int nobs, int nthreads) {
    if (nobs < 0 || nobs > MAX_OBSTACLE_NUM) return nullptr;  // :90-95
    RefArmtd* F = new RefArmtd();
    const int n = NUM_FACTORS * NUM_TIME_STEPS;
    F->q0.assign(q0, q0 + NUM_FACTORS);
    F->qd0.assign(qd0, qd0 + NUM_FACTORS);
    F->q_des.assign(q_des, q_des + NUM_FACTORS);
    F->k_range.assign(k_range, k_range + NUM_FACTORS);
    for (int a = 0; a < 6; a++) F->jrs[a].assign(jrs + size_t(a) * n, jrs + size_t(a + 1) * n);
    F->nobs = nobs;
    std::memset(F->obstacles, 0, sizeof(F->obstacles));
    std::memcpy(F->obstacles, obstacles, sizeof(double) * nobs * (MAX_OBSTACLE_GENERATOR_NUM + 1) * 3);
    if (nthreads > 0) omp_set_num_threads(nthreads);
    try {
        F->traj = ConstantAccelerationCurve(F->q0.data(), F->qd0.data(), F->jrs[0].data(), F->jrs[1].data(), F->jrs[2].data(),
                                            F->jrs[3].data(), F->jrs[4].data(), F->jrs[5].data(), F->k_range.data());  // :110-113
        F->O = new Obstacles(F->obstacles, nobs);                                                                    // :115
        F->kd = new KinematicsDynamics(&F->traj);                                                                     // :117
        F->gens = new Eigen::Matrix<double, 3, 3 + 3>[NUM_TIME_STEPS * NUM_JOINTS];
        int t = 0;
#pragma omp parallel for shared(F) private(t) schedule(dynamic, 1)
        for (t = 0; t < NUM_TIME_STEPS; t++) F->traj.makePolyZono(t);  // :129-132
#pragma omp parallel for shared(F) private(t) schedule(dynamic)
        for (t = 0; t < NUM_TIME_STEPS; t++) {  // :140-149
            F->kd->fk(t);
            for (int i = 0; i < NUM_JOINTS; i++) F->gens[t * NUM_JOINTS + i] = F->kd->links(i, t).reduce_link_PZ();
        }
        F->O->initializeHyperPlane(F->gens);  // :157
        if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) throw 2;
        F->nlp = new armtd_NLP();
        F->nlp->set_parameters(F->q_des.data(), &F->traj, F->kd, F->O);  // :171-172
    } catch (...) {
        delete F;
        return nullptr;
    }
    return F;
}
void refarmtd_destroy(void* h) { delete static_cast<RefArmtd*>(h); }
int refarmtd_num_constraints(void* h) { return static_cast<RefArmtd*>(h)->nlp->constraint_number; }
void refarmtd_bounds(void* h, double* x_l, double* x_u, double* g_l, double* g_u) {
    RefArmtd* F = static_cast<RefArmtd*>(h);
    F->nlp->get_bounds_info(NUM_FACTORS, x_l, x_u, F->nlp->constraint_number, g_l, g_u);
}
void refarmtd_cost(void* h, const double* k, double* f, double* grad) {
    RefArmtd* F = static_cast<RefArmtd*>(h);
    F->nlp->eval_f(NUM_FACTORS, k, true, *f);
    F->nlp->eval_grad_f(NUM_FACTORS, k, true, grad);
}
int refarmtd_eval_g(void* h, const double* k, double* g) {
    RefArmtd* F = static_cast<RefArmtd*>(h);
    F->nlp->eval_g(NUM_FACTORS, k, true, F->nlp->constraint_number, g);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
// values must be zero-filled by the caller: returnJointStateExtremumGradient clears only the first 4*7*7 BYTES of its 28 rows
// (KPA/Trajectory.cu:262) and writes the diagonal entries, so the other entries keep what the buffer held
int refarmtd_eval_jac_g(void* h, const double* k, double* values) {
    RefArmtd* F = static_cast<RefArmtd*>(h);
    const int m = F->nlp->constraint_number;
    F->nlp->eval_jac_g(NUM_FACTORS, k, true, m, m * NUM_FACTORS, nullptr, nullptr, values);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
int refarmtd_finalize(void* h, const double* k, const double* g, double obj) {
    RefArmtd* F = static_cast<RefArmtd*>(h);
    F->nlp->finalize_solution(Ipopt::SUCCESS, NUM_FACTORS, k, nullptr, nullptr, F->nlp->constraint_number, g, nullptr, obj, nullptr,
                              nullptr);
    return F->nlp->feasible ? 1 : 0;
}
void refarmtd_link_sliced_center(void* h, double* out) {
    RefArmtd* F = static_cast<RefArmtd*>(h);
    for (int i = 0; i < NUM_TIME_STEPS * NUM_JOINTS; i++)
        for (int e = 0; e < 3; e++) out[i * 3 + e] = F->nlp->link_sliced_center[i](e);
}
// link generator matrices as armtd_main.cu writes them (:245-256): out[(t*NJ + l)*18], column-major 3x6
void refarmtd_link_gens(void* h, double* out) {
    RefArmtd* F = static_cast<RefArmtd*>(h);
    for (int i = 0; i < NUM_TIME_STEPS * NUM_JOINTS; i++)
        for (int c = 0; c < 6; c++)
            for (int r = 0; r < 3; r++) out[i * 18 + c * 3 + r] = F->gens[i](r, c);
}
// k-only monomials of the link reach sets after reduce_link_PZ, for table-level comparisons: per (t, l) the number of monomials,
// then hash and the 3 coefficients; returns the number of doubles written (at most cap)
int refarmtd_link_tables(void* h, double* out, int cap) {
    RefArmtd* F = static_cast<RefArmtd*>(h);
    int k = 0;
    for (int t = 0; t < NUM_TIME_STEPS; t++)
        for (int l = 0; l < NUM_JOINTS; l++) {
            const PZsparse& z = F->kd->links(l, t);
            if (k + 4 + int(z.polynomial.size()) * 4 > cap) return -1;
            out[k++] = double(z.polynomial.size());
            for (int e = 0; e < 3; e++) out[k++] = z.center(e, 0);
            for (const Monomial& mono : z.polynomial) {
                out[k++] = double(mono.degree);
                for (int e = 0; e < 3; e++) out[k++] = mono.coeff(e, 0);
            }
        }
    return k;
}
}
