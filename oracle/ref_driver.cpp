// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Thin C driver around the REFERENCE's own classes (PZsparse, BezierCurve, KinematicsDynamics), compiled from
// the reference's source files where they lie under /root/reference (oracle/Makefile.ref; nothing is copied
// into this repository) against the stand-in headers of oracle/ref_shim/ (Eigen, Boost.Interval and the CUDA
// runtime are not installed here).  The result, oracle/_ref/libarmour_ref.so, is used ONLY to pin the restated
// oracle (tests/test_oracle_vs_reference.py, tools/make_golden.py).
//
// The driver itself restates only the orchestration of main(): KPR/armour_main.cu:96-142 (JRS, FK,
// reduce_link_PZ, nominal + interval RNEA, disturbance, reduce) and :172-201 (robust-input radius), then the
// slicing loops of armtd_NLP::eval_g / eval_jac_g (KPR/NLPclass.cu:304-315, 376-391).  The collision rows need
// the reference's CUDA kernels (KPR/CollisionChecking.cu) and are not produced here.
#include "ref_problem.h"

extern "C" {

int ref_num_joints() { return NUM_JOINTS; }
int ref_num_time_steps() { return NUM_TIME_STEPS; }

void* ref_build(const double* q0, const double* qd0, const double* qdd0, int nthreads) {
    RefProblem* P = new RefProblem();
    if (!ref_problem_build(P, q0, qd0, qdd0, nthreads)) {
        delete P;
        return nullptr;
    }
    return P;
}

void ref_destroy(void* h) { delete static_cast<RefProblem*>(h); }

// Same neutral table layout as orc_export_reachsets (oracle/oracle_capi.cpp) and armour_export_reachsets.
// Returns the largest monomial count, or its negative if a capacity is exceeded.
int ref_export(void* h, int cap_link, int* link_n, double* link_center, unsigned long long* link_key, double* link_coeff,
               int cap_u, int* u_n, double* u_center, unsigned long long* u_key, double* u_coeff, double* u_radius,
               double* torque_radius, double* link_gens) {
    RefProblem* P = static_cast<RefProblem*>(h);
    int mx = 0;
    for (int t = 0; t < NUM_TIME_STEPS; t++) {
        for (int l = 0; l < NUM_JOINTS; l++) {
            const PZsparse& z = P->kd.links(l, t);
            const int i = t * NUM_JOINTS + l, n = int(z.polynomial.size());
            mx = std::max(mx, n);
            if (n > cap_link) return -n;
            link_n[i] = n;
            for (int e = 0; e < 3; e++) link_center[i * 3 + e] = z.center(e);
            for (int m = 0; m < n; m++) {
                link_key[i * cap_link + m] = z.polynomial[m].degree;
                for (int e = 0; e < 3; e++) link_coeff[(i * cap_link + m) * 3 + e] = z.polynomial[m].coeff(e);
            }
            for (int e = 0; e < 18; e++) link_gens[i * 18 + e] = P->link_gens[i](e);
        }
        for (int j = 0; j < NUM_FACTORS; j++) {
            const PZsparse& z = P->kd.u_nom(j, t);
            const int i = t * NUM_FACTORS + j, n = int(z.polynomial.size());
            mx = std::max(mx, n);
            if (n > cap_u) return -n;
            u_n[i] = n;
            u_center[i] = z.center(0);
            u_radius[i] = z.independent(0);
            for (int m = 0; m < n; m++) {
                u_key[i * cap_u + m] = z.polynomial[m].degree;
                u_coeff[i * cap_u + m] = z.polynomial[m].coeff(0);
            }
            torque_radius[j * NUM_TIME_STEPS + t] = P->torque_radius(j, t);
        }
    }
    return mx;
}

// Slices at k with the reference's PZsparse::slice overloads and the Bezier extremum rows
// (KPR/NLPclass.cu:304-320, 376-391): g_torque[T*7], jac_torque[T*7*7], link_c[T*NJ*3], dlink_c[T*NJ*7*3],
// bez[28], dbez[28*7].
void ref_slice(void* h, const double* k, double* g_torque, double* jac_torque, double* link_c, double* dlink_c,
               double* bez, double* dbez) {
    RefProblem* P = static_cast<RefProblem*>(h);
    for (int t = 0; t < NUM_TIME_STEPS; t++) {
        for (int j = 0; j < NUM_FACTORS; j++) {
            MatrixXInt res = P->kd.u_nom(j, t).slice(k);
            g_torque[t * NUM_FACTORS + j] = getCenter(res(0));
            P->kd.u_nom(j, t).slice(jac_torque + (t * NUM_FACTORS + j) * NUM_FACTORS, k);
        }
        for (int l = 0; l < NUM_JOINTS; l++) {
            MatrixXInt res = P->kd.links(l, t).slice(k);
            Eigen::MatrixXd c = getCenter(res);
            for (int e = 0; e < 3; e++) link_c[(t * NUM_JOINTS + l) * 3 + e] = c(e);
            Eigen::Vector3d grad[NUM_FACTORS];
            P->kd.links(l, t).slice(grad, k);
            for (int v = 0; v < NUM_FACTORS; v++)
                for (int e = 0; e < 3; e++) dlink_c[((t * NUM_JOINTS + l) * NUM_FACTORS + v) * 3 + e] = grad[v](e);
        }
    }
    P->traj.returnJointPositionExtremum(bez, k);
    P->traj.returnJointVelocityExtremum(bez + NUM_FACTORS * 2, k);
    std::memset(dbez, 0, sizeof(double) * 4 * NUM_FACTORS * NUM_FACTORS);
    P->traj.returnJointPositionExtremumGradient(dbez, k);
    P->traj.returnJointVelocityExtremumGradient(dbez + NUM_FACTORS * 2 * NUM_FACTORS, k);
}

// JRS pieces of joint i, interval t for diagnostics: centre and the monomial coefficients of cos/sin
int ref_jrs(void* h, int i, int t, double* out /* [cos c, n, coeffs..8][sin ...] 20 doubles */) {
    RefProblem* P = static_cast<RefProblem*>(h);
    const PZsparse* z[2] = {&P->traj.cos_q_des(i, t), &P->traj.sin_q_des(i, t)};
    for (int s = 0; s < 2; s++) {
        double* o = out + s * 10;
        std::memset(o, 0, 10 * sizeof(double));
        o[0] = z[s]->center(0);
        o[1] = double(z[s]->polynomial.size());
        o[2] = z[s]->independent(0);
        for (size_t m = 0; m < z[s]->polynomial.size() && m < 3; m++) o[3 + m] = z[s]->polynomial[m].coeff(0);
    }
    return 0;
}

}  // extern "C"
