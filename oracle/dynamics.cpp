// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned: bit-identical to the reference's own sources (oracle/_ref, tests/test_oracle_pinned.py, tests/golden/ref).
// See dynamics.h.  Line references are to the reference file KPR/Dynamics.cu.
#include "dynamics.h"

#include <cstdlib>

namespace orc {

KinematicsDynamics::KinematicsDynamics(BezierCurve* tr) : model(tr->model), traj(tr), T(tr->T) {  // :6-67
    const RobotModel& m = *model;
    const int nj = m.num_joints;
    mass_nominal.resize(nj);
    mass_uncertain.resize(nj);
    I_nominal.resize(nj);
    I_uncertain.resize(nj);
    links.resize(nj * T);
    u_nom.resize(NF * T);
    u_nom_int.resize(NF * T);
    for (int i = 0; i < nj; i++) {
        mass_nominal[i] = PZ::matrix(1, 1, &m.mass[i]);
        mass_uncertain[i] = PZ::matrix_uncertain(1, 1, &m.mass[i], m.mass_uncertainty);
        // inertia_matrix(j) = inertia[i*9 + j]: linear (column-major) index, :36-38
        I_nominal[i] = PZ::matrix(3, 3, &m.inertia[i * 9]);
        I_uncertain[i] = PZ::matrix_uncertain(3, 3, &m.inertia[i * 9], m.inertia_uncertainty);
    }
    for (int i = 0; i < nj; i++) {  // link box zonotopes, :51-66
        PZ comp[3];
        for (int j = 0; j < 3; j++) {
            const uint64_t h = var_hash(NF * (j + 1));  // qde_0 / qdae_0 / qddae_0 slots carry x / y / z
            const double g = m.link_zonotope_generators[i][j];
            comp[j] = PZ::scalar_poly(m.link_zonotope_center[i][j], &g, &h, 1);
        }
        const PZ link = stack3(comp[0], comp[1], comp[2]);
        for (int t = 0; t < T; t++) links[i * T + t] = link;
    }
}

void KinematicsDynamics::fk(int t) {  // :69-81
    const RobotModel& m = *model;
    PZ FK_R = PZ::rpy(0, 0, 0);
    PZ FK_T(3, 1);
    for (int i = 0; i < m.num_joints; i++) {
        const PZ P = PZ::matrix(3, 1, &m.trans[3 * i]);
        FK_T = FK_T + FK_R * P;
        FK_R = FK_R * traj->R[i * T + t];
        links[i * T + t] = FK_R * links[i * T + t] + FK_T;
        if (probe) {
            probe("FK_R", i, FK_R);
            probe("FK_T", i, FK_T);
            probe("link", i, links[i * T + t]);
        }
    }
}

void KinematicsDynamics::rnea(int t, const std::vector<PZ>& mass_arr, const std::vector<PZ>& I_arr,
                              std::vector<PZ>& u) {  // :83-181
    const RobotModel& m = *model;
    const int nj = m.num_joints;
    PZ w(3, 1), wdot(3, 1), w_aux(3, 1), linear_acc(3, 1);
    std::vector<PZ> F(nj), N(nj);
    linear_acc.center[2] = m.gravity;

    for (int i = 0; i < nj; i++) {
        const double* p = &m.trans[3 * i];
        const double* c = &m.com[3 * i];
        const PZ& Rt = traj->R_t[i * T + t];
        linear_acc = Rt * ((linear_acc + cross_pz_mat(wdot, p)) + cross_pz_pz(w, cross_pz_mat(w_aux, p)));
        if (m.axes[i] != 0) {
            const int ax = std::abs(m.axes[i]) - 1;
            const PZ& qd = traj->qd_des[i * T + t];
            w = Rt * w;
            w.add_one_dim(qd, ax, 0);
            w_aux = Rt * w_aux;
            wdot = Rt * wdot;
            PZ temp(3, 1);
            temp.add_one_dim(qd, ax, 0);
            wdot = wdot + cross_pz_pz(w_aux, temp);
            wdot.add_one_dim(traj->qdda_des[i * T + t], ax, 0);
            w_aux.add_one_dim(traj->qda_des[i * T + t], ax, 0);
        } else {
            w = Rt * w;
            w_aux = Rt * w_aux;
            wdot = Rt * wdot;
        }
        F[i] = mass_arr[i] * ((linear_acc + cross_pz_mat(wdot, c)) + cross_pz_pz(w, cross_pz_mat(w_aux, c)));
        N[i] = I_arr[i] * wdot + cross_pz_pz(w_aux, I_arr[i] * w);
        if (probe) {
            probe("linear_acc", i, linear_acc);
            probe("w", i, w);
            probe("w_aux", i, w_aux);
            probe("wdot", i, wdot);
            probe("F", i, F[i]);
            probe("N", i, N[i]);
        }
    }

    PZ f(3, 1), n(3, 1);
    for (int i = nj - 1; i >= 0; i--) {
        const PZ& Rn = traj->R[(i + 1) * T + t];
        n = ((N[i] + Rn * n) + cross_mat_pz(&m.com[3 * i], F[i])) + cross_mat_pz(&m.trans[3 * (i + 1)], Rn * f);
        f = Rn * f + F[i];
        if (probe) {
            probe("n", i, n);
            probe("f", i, f);
        }
        if (m.axes[i] != 0) {
            const int ax = std::abs(m.axes[i]) - 1;
            PZ ui = n.elem(ax, 0);
            ui = ui + scale(m.armature[i], traj->qdda_des[i * T + t]);
            ui = ui + scale(m.damping[i], traj->qd_des[i * T + t]);
            u[i * T + t] = ui;
            if (probe) probe("u", i, ui);
        }
    }
}

}  // namespace orc
