// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned against the reference's own sources: reach-set build, slices and Bezier rows through oracle/_ref/libarmour_ref.so (tests/test_oracle_pinned.py); collision rows, bounds, cost and verdict through the reference's CUDA kernels and armtd_NLP compiled by nvcc (oracle/_ref/libarmour_ref_cuda.so, run on a B200, frozen as tests/golden/refcuda/, tests/test_refcuda_golden.py).
// Plain C entry points over orc::Problem so tests / bench.py can drive the oracle with ctypes.
#include <cstring>

#include "planner.h"

using orc::NF;
using orc::Problem;

extern "C" {

void* orc_create(int model_id, int num_time_steps, double simplify_threshold, const double* k_range, int max_obstacles,
                 double mass_uncertainty, double inertia_uncertainty) {
    orc::PlannerParams p;
    if (num_time_steps > 0) p.num_time_steps = num_time_steps;
    if (simplify_threshold > 0) p.simplify_threshold = simplify_threshold;
    if (k_range)
        for (int i = 0; i < NF; i++) p.k_range[i] = k_range[i];
    if (max_obstacles > 0) p.max_obstacles = max_obstacles;
    Problem* P = new Problem(model_id, p);
    if (mass_uncertainty >= 0) P->model.mass_uncertainty = mass_uncertainty;
    if (inertia_uncertainty >= 0) P->model.inertia_uncertainty = inertia_uncertainty;
    return P;
}
void orc_destroy(void* h) { delete static_cast<Problem*>(h); }

int orc_build(void* h, const double* q0, const double* qd0, const double* qdd0, const double* obstacles, int nobs,
              int nthreads) {
    try {
        static_cast<Problem*>(h)->build(q0, qd0, qdd0, obstacles, nobs, nthreads);
    } catch (...) {
        return -1;
    }
    return 0;
}
int orc_num_constraints(void* h) { return static_cast<Problem*>(h)->num_constraints(); }
int orc_num_joints(void* h) { return static_cast<Problem*>(h)->NJ; }
int orc_num_time_steps(void* h) { return static_cast<Problem*>(h)->T; }
double orc_build_ms(void* h) { return static_cast<Problem*>(h)->build_ms; }
void orc_eval_g(void* h, const double* k, double* g) { static_cast<Problem*>(h)->eval_g(k, g); }
void orc_eval_jac_g(void* h, const double* k, double* v) { static_cast<Problem*>(h)->eval_jac_g(k, v); }
void orc_bounds(void* h, double* gl, double* gu) { static_cast<Problem*>(h)->bounds(gl, gu); }
int orc_verdict(void* h, const double* g, int* first) { return static_cast<Problem*>(h)->verdict(g, first); }
double orc_cost(void* h, const double* q_des, const double* k) { return static_cast<Problem*>(h)->cost(q_des, k); }
void orc_cost_grad(void* h, const double* q_des, const double* k, double* grad) {
    static_cast<Problem*>(h)->cost_grad(q_des, k, grad);
}
void orc_get_torque_radius(void* h, double* out) {  // [j*T + t]
    Problem* P = static_cast<Problem*>(h);
    std::memcpy(out, P->torque_radius.data(), sizeof(double) * P->torque_radius.size());
}
void orc_get_link_gens(void* h, double* out) {  // [t*NJ + l][18] column-major 3x6
    Problem* P = static_cast<Problem*>(h);
    std::memcpy(out, P->link_gens.data(), sizeof(double) * P->link_gens.size());
}
void orc_get_link_sliced_center(void* h, double* out) {  // [t*NJ + l][3]
    Problem* P = static_cast<Problem*>(h);
    std::memcpy(out, P->link_sliced_center.data(), sizeof(double) * P->link_sliced_center.size());
}
void orc_get_hyperplanes(void* h, double* A, double* d, double* delta) {
    Problem* P = static_cast<Problem*>(h);
    std::memcpy(A, P->A.data(), sizeof(double) * P->A.size());
    std::memcpy(d, P->d.data(), sizeof(double) * P->d.size());
    std::memcpy(delta, P->delta.data(), sizeof(double) * P->delta.size());
}
void orc_get_jrs(void* h, double* out) {  // [i*T + t][13], field order of orc::JrsDump
    Problem* P = static_cast<Problem*>(h);
    std::memcpy(out, P->traj->dump.data(), sizeof(orc::JrsDump) * P->traj->dump.size());
}
void orc_get_stats(void* h, unsigned long long* out) {
    const orc::Stats& s = static_cast<Problem*>(h)->stats;
    out[0] = s.n_simplify;
    out[1] = s.n_mul;
    out[2] = s.n_pairs;
    out[3] = s.flops;
    out[4] = s.terms_sorted;
    out[5] = s.max_terms;
    out[6] = s.max_monos;
    out[7] = s.near_threshold;
}

// Re-run FK + nominal RNEA of interval t on scratch copies with the op trace on; returns the number of
// uint32 written: 7 per simplify() = {is_product, n1, n2, terms_before, monomials_after, coeff_size, unique_keys}.
int orc_trace_interval(void* h, int t, unsigned int* out, int cap) {
    Problem* P = static_cast<Problem*>(h);
    orc::tls_threshold() = P->params.simplify_threshold;
    std::vector<uint32_t> tr;
    orc::KinematicsDynamics kd(P->traj.get());
    orc::tls_trace() = &tr;
    kd.fk(t);
    tr.push_back(0xFFFFFFFFu);  // separator: FK done
    for (int i = 0; i < 6; i++) tr.push_back(0);
    kd.rnea_nominal(t);
    orc::tls_trace() = nullptr;
    const int n = int(std::min<size_t>(tr.size(), size_t(cap)));
    std::memcpy(out, tr.data(), sizeof(uint32_t) * n);
    return int(tr.size());
}

// Sizes of the named intermediates of interval t (nominal RNEA): out[(name_id*8 + joint)] = monomial count,
// name ids: 0 FK_R 1 FK_T 2 link 3 linear_acc 4 w 5 w_aux 6 wdot 7 F 8 N 9 n 10 f 11 u
void orc_probe_sizes(void* h, int t, int* out) {
    Problem* P = static_cast<Problem*>(h);
    orc::tls_threshold() = P->params.simplify_threshold;
    static const char* names[12] = {"FK_R", "FK_T", "link", "linear_acc", "w", "w_aux", "wdot", "F", "N", "n", "f", "u"};
    for (int i = 0; i < 12 * 8; i++) out[i] = 0;
    orc::KinematicsDynamics kd(P->traj.get());
    kd.probe = [&](const char* nm, int joint, const orc::PZ& z) {
        for (int i = 0; i < 12; i++)
            if (!std::strcmp(nm, names[i])) out[i * 8 + joint] = int(z.poly.size());
    };
    kd.fk(t);
    kd.rnea_nominal(t);
}

// k-only reach-set tables after reduce / reduce_link_PZ, in a neutral padded layout:
//   links : n[t*NJ+l], center[(t*NJ+l)*3+e], hash[(t*NJ+l)*cap+m], coeff[((t*NJ+l)*cap+m)*3+e]
//   torque: n[t*NF+j], center[t*NF+j],       hash[(t*NF+j)*cap+m], coeff[(t*NF+j)*cap+m], radius[t*NF+j]
// returns the largest monomial count seen, or -(that count) if a capacity was exceeded.
int orc_export_reachsets(void* h, int cap_link, int* nl, double* cl, unsigned long long* hl, double* gl, int cap_u,
                         int* nu, double* cu, unsigned long long* hu, double* gu, double* ru) {
    Problem* P = static_cast<Problem*>(h);
    const int T = P->T, NJ = P->NJ;
    int mx = 0;
    bool over = false;
    for (int t = 0; t < T; t++) {
        for (int l = 0; l < NJ; l++) {
            const orc::PZ& z = P->kd->links[l * T + t];
            const int idx = t * NJ + l, n = int(z.poly.size());
            mx = std::max(mx, n);
            nl[idx] = n;
            for (int e = 0; e < 3; e++) cl[idx * 3 + e] = z.center[e];
            if (n > cap_link) { over = true; continue; }
            for (int m = 0; m < n; m++) {
                hl[size_t(idx) * cap_link + m] = z.poly[m].degree;
                for (int e = 0; e < 3; e++) gl[(size_t(idx) * cap_link + m) * 3 + e] = z.poly[m].c[e];
            }
        }
        for (int j = 0; j < NF; j++) {
            const orc::PZ& z = P->kd->u_nom[j * T + t];
            const int idx = t * NF + j, n = int(z.poly.size());
            mx = std::max(mx, n);
            nu[idx] = n;
            cu[idx] = z.center[0];
            ru[idx] = z.indep[0];
            if (n > cap_u) { over = true; continue; }
            for (int m = 0; m < n; m++) {
                hu[size_t(idx) * cap_u + m] = z.poly[m].degree;
                gu[size_t(idx) * cap_u + m] = z.poly[m].c[0];
            }
        }
    }
    return over ? -mx : mx;
}

}  // extern "C"
