// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// C driver around the REFERENCE's complete planner path, CUDA kernels included: the reference's own
// PZsparse.cu, Trajectory.cu, Dynamics.cu, CollisionChecking.cu and NLPclass.cu are compiled by nvcc from where they
// lie under /root/reference (oracle/Makefile.ref, target _ref/libarmour_ref_cuda.so; nothing is copied into this
// repository) against the stand-in headers of oracle/ref_shim/ (Eigen with __host__ __device__ fixed-size
// matrices, Boost.Interval, and the few Ipopt declarations armtd_NLP derives from).  The library needs a GPU:
// tools/make_golden_collision.py runs it on a B200 through gpurun and freezes its outputs as
// tests/golden/refcuda/*.npz, which pin the collision rows, the bounds and the verdict of the restated oracle
// and of the CUDA product path against the reference ITSELF (not against a restatement).
//
// The driver restates only what main() does around those classes (KPR/armour_main.cu:86-216, 230-236: Obstacles
// ctor, reach sets, robust-input radius, initializeHyperPlane, armtd_NLP::set_parameters) and then calls the
// armtd_NLP members the way Ipopt would.  There is no solver here.
#include "NLPclass.h"  // the reference's header: Dynamics.h, CollisionChecking.h, armtd_NLP

#include "ref_problem.h"

namespace {
struct RefFull {
    RefProblem P;
    double obstacles[MAX_OBSTACLE_NUM * (MAX_OBSTACLE_GENERATOR_NUM + 1) * 3];
    int nobs = 0;
    Obstacles* O = nullptr;
    armtd_NLP* nlp = nullptr;
    Eigen::Matrix<double, 3, 3 + 3>* gens = nullptr;  // [t*NUM_JOINTS + l] as main() holds them (armour_main.cu:113)
    ~RefFull() {
        delete nlp;
        delete O;
        delete[] gens;
    }
};
}  // namespace

extern "C" {

int reffull_cuda_devices() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// q0/qd0/qdd0/q_des [7], obstacles [nobs*12].  Returns nullptr when the reference threw or CUDA failed.
void* reffull_build(const double* q0, const double* qd0, const double* qdd0, const double* q_des, const double* obstacles,
                    int nobs, int nthreads) {
    if (nobs < 0 || nobs > MAX_OBSTACLE_NUM) return nullptr;  // armour_main.cu:66-71 (the reference throws)
    RefFull* F = new RefFull();
    F->nobs = nobs;
    std::memcpy(F->obstacles, obstacles, sizeof(double) * nobs * (MAX_OBSTACLE_GENERATOR_NUM + 1) * 3);
    try {
        F->O = new Obstacles(F->obstacles, nobs);                           // armour_main.cu:86
        if (!ref_problem_build(&F->P, q0, qd0, qdd0, nthreads)) throw 1;    // :96-201
        F->gens = new Eigen::Matrix<double, 3, 3 + 3>[NUM_TIME_STEPS * NUM_JOINTS];
        for (int i = 0; i < NUM_TIME_STEPS * NUM_JOINTS; i++) F->gens[i] = F->P.link_gens[i];
        F->O->initializeHyperPlane(F->gens);                                // :208
        if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) throw 2;
        Eigen::VectorXd qd(NUM_FACTORS);
        for (int i = 0; i < NUM_FACTORS; i++) qd(i) = q_des[i];
        F->nlp = new armtd_NLP();
        F->nlp->set_parameters(qd, 0.5, &F->P.traj, &F->P.kd, &F->P.torque_radius, F->O);  // :230-236, t_plan :80
    } catch (...) {
        delete F;
        return nullptr;
    }
    return F;
}

void reffull_destroy(void* h) { delete static_cast<RefFull*>(h); }

int reffull_num_constraints(void* h) {
    Ipopt::Index n, m, nnz_j, nnz_h;
    Ipopt::TNLP::IndexStyleEnum st;
    static_cast<RefFull*>(h)->nlp->get_nlp_info(n, m, nnz_j, nnz_h, st);
    return m;
}

void reffull_bounds(void* h, double* x_l, double* x_u, double* g_l, double* g_u) {
    RefFull* F = static_cast<RefFull*>(h);
    F->nlp->get_bounds_info(NUM_FACTORS, x_l, x_u, F->nlp->constraint_number, g_l, g_u);
}

void reffull_cost(void* h, const double* k, double* f, double* grad) {
    RefFull* F = static_cast<RefFull*>(h);
    F->nlp->eval_f(NUM_FACTORS, k, true, *f);
    F->nlp->eval_grad_f(NUM_FACTORS, k, true, grad);
}

int reffull_eval_g(void* h, const double* k, double* g) {
    RefFull* F = static_cast<RefFull*>(h);
    F->nlp->eval_g(NUM_FACTORS, k, true, F->nlp->constraint_number, g);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int reffull_eval_jac_g(void* h, const double* k, double* values) {
    RefFull* F = static_cast<RefFull*>(h);
    const int m = F->nlp->constraint_number;
    F->nlp->eval_jac_g(NUM_FACTORS, k, true, m, m * NUM_FACTORS, nullptr, nullptr, values);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// finalize_solution's verdict on (k, g) (KPR/NLPclass.cu:426-537).  It prints the violated row to stdout like the
// reference does.
int reffull_finalize(void* h, const double* k, const double* g, double obj) {
    RefFull* F = static_cast<RefFull*>(h);
    F->nlp->finalize_solution(Ipopt::SUCCESS, NUM_FACTORS, k, nullptr, nullptr, F->nlp->constraint_number, g, nullptr, obj,
                              nullptr, nullptr);
    return F->nlp->feasible ? 1 : 0;
}

// the sliced link centres of the last eval (armtd_NLP::link_sliced_center, KPR/NLPclass.h:150): out[(t*NJ+l)*3]
void reffull_link_sliced_center(void* h, double* out) {
    RefFull* F = static_cast<RefFull*>(h);
    for (int i = 0; i < NUM_TIME_STEPS * NUM_JOINTS; i++)
        for (int e = 0; e < 3; e++) out[i * 3 + e] = F->nlp->link_sliced_center[i](e);
}

// The stored half-spaces of the reference, device -> host, for diagnosing argmax differences:
// A[(((t*NJ+l)*nobs+o)*36+p)*3], d and delta [((t*NJ+l)*nobs+o)*36+p]  (KPR/CollisionChecking.cu:230-241)
int reffull_hyperplanes(void* h, double* A, double* d, double* delta) {
    RefFull* F = static_cast<RefFull*>(h);
    const size_t n = size_t(NUM_TIME_STEPS) * NUM_JOINTS * F->nobs * COMB_NUM;
    if (cudaMemcpy(A, F->O->dev_A, n * sizeof(Eigen::Vector3d), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (cudaMemcpy(d, F->O->dev_d, n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (cudaMemcpy(delta, F->O->dev_delta, n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return 0;
}

void* reffull_problem(void* h) { return &static_cast<RefFull*>(h)->P; }  // for ref_export / ref_slice of ref_driver.cpp

}  // extern "C"
