// ARMTD comparison planner (reference directory kinova_planner_realtime_armtd_comparison = "KPA") on the
// kernels of the main path.  KPA differs from the main planner in two places only: the trajectory class
// (KPA/Trajectory.cu: constant-acceleration trajectories whose cos / sin reach sets come from an OFFLINE table handed in by the
// caller) and the NLP (KPA/NLPclass.cu: no torque rows, other joint-limit rows, other cost); its PZ arithmetic, forward
// kinematics, reduce_link_PZ and collision kernels are copies of the main planner's.  So the build is k_reachsets with the
// joint reachable set imported (Batch::jrs_ext, csrc/k1_reachsets.cuh: jrs_joint) followed by k_hyperplanes, an evaluation is
// k_constraints followed by the kernel below, which lays the collision rows out as KPA does and adds its joint-limit rows.
// KPA uses 100 time steps (KPA/Parameters.h:17); the constraint kernels work on chunks of 8 intervals, so the context runs 104
// with the last interval repeated and the padding rows are dropped here.
#pragma once
#include <cuda_runtime.h>

#include "layout.h"

namespace armour {

constexpr int ARMTD_T = 100;       // NUM_TIME_STEPS of KPA
constexpr int ARMTD_T_PADDED = 104;

struct ArmtdParams {
    double q0[NF], qd0[NF], k_range[NF];
};

// ConstantAccelerationCurve::returnJointStateExtremum and ...Gradient for one joint (KPA/Trajectory.cu:88-384):
// ext / grad = {q_min, q_max, qd_min, qd_max}; the gradients are with respect to k_range * k, as the reference writes them
__host__ __device__ inline void armtd_joint_extremum(double q0, double qd0, double k_actual, double* ext, double* grad) {
    const double t_move = 0.5, t_total = 1.0, t_to_stop = t_total - t_move;
    const double q_peak = q0 + qd0 * t_move + k_actual * t_move * t_move * 0.5;
    const double q_dot_peak = qd0 + k_actual * t_move;
    const double q_ddot_to_stop = -q_dot_peak / t_to_stop;
    const double q_stop = q_peak + q_dot_peak * t_to_stop + 0.5 * q_ddot_to_stop * t_to_stop * t_to_stop;
    const double t_mm = -qd0 / k_actual;  // interior extremum of the first phase; inf / nan at k = 0 fails both tests below
    double q_lo, q_hi, g_lo, g_hi;
    if (q_peak >= q0) {
        q_lo = q0; q_hi = q_peak; g_lo = 0; g_hi = 0.5 * t_move * t_move;
    } else {
        q_lo = q_peak; q_hi = q0; g_lo = 0.5 * t_move * t_move; g_hi = 0;
    }
    double q_min_p = q_lo, q_max_p = q_hi, gq_min_p = g_lo, gq_max_p = g_hi;
    if (t_mm > 0 && t_mm < t_move) {
        const double q_int = q0 + qd0 * t_mm + 0.5 * k_actual * t_mm * t_mm;
        const double g_int = (0.5 * qd0 * qd0) / (k_actual * k_actual);
        if (k_actual >= 0) {
            q_min_p = q_int; gq_min_p = g_int;
        } else {
            q_max_p = q_int; gq_max_p = g_int;
        }
    }
    double v_min_p, v_max_p, gv_min_p, gv_max_p;
    if (q_dot_peak >= qd0) {
        v_min_p = qd0; v_max_p = q_dot_peak; gv_min_p = 0; gv_max_p = t_move;
    } else {
        v_min_p = q_dot_peak; v_max_p = qd0; gv_min_p = t_move; gv_max_p = 0;
    }
    double q_min_s, q_max_s, gq_min_s, gq_max_s;
    if (q_stop >= q_peak) {
        q_min_s = q_peak; q_max_s = q_stop;
        gq_min_s = 0.5 * t_move * t_move; gq_max_s = 0.5 * t_move * t_move + 0.5 * t_move * t_to_stop;
    } else {
        q_min_s = q_stop; q_max_s = q_peak;
        gq_min_s = 0.5 * t_move * t_move + 0.5 * t_move * t_to_stop; gq_max_s = 0.5 * t_move * t_move;
    }
    double v_min_s, v_max_s, gv_min_s, gv_max_s;
    if (q_dot_peak >= 0) {
        v_min_s = 0; v_max_s = q_dot_peak; gv_min_s = 0; gv_max_s = t_move;
    } else {
        v_min_s = q_dot_peak; v_max_s = 0; gv_min_s = t_move; gv_max_s = 0;
    }
    const bool a = q_min_p <= q_min_s, b = q_max_p >= q_max_s, c = v_min_p <= v_min_s, d = v_max_p >= v_max_s;
    ext[0] = a ? q_min_p : q_min_s;
    ext[1] = b ? q_max_p : q_max_s;
    ext[2] = c ? v_min_p : v_min_s;
    ext[3] = d ? v_max_p : v_max_s;
    grad[0] = a ? gq_min_p : gq_min_s;
    grad[1] = b ? gq_max_p : gq_max_s;
    grad[2] = c ? gv_min_p : gv_min_s;
    grad[3] = d ? gv_max_p : gv_max_s;
}

// g_full / j_full: one problem in the main planner's layout with Tp = 104 intervals (torque rows, collision rows
// (l * Tp + t) * O + o, Bezier rows); g / jac: KPA's layout, collision rows (l * 100 + t) * O + o then 4 * NF joint-limit rows
// (KPA/NLPclass.cu:43-44, 248-336).  The joint-limit rows of the Jacobian are diagonal: the reference writes only the diagonal
// entries and clears 4*NF*NF BYTES of those rows (KPA/Trajectory.cu:262), i.e. it relies on a zero-filled buffer.
__global__ void __launch_bounds__(256)
k_armtd_assemble(ArmtdParams A, int NJ, int O, const double* __restrict__ k, const double* __restrict__ g_full,
                 const double* __restrict__ j_full, double* __restrict__ g, double* __restrict__ jac) {
    const int rows = NJ * ARMTD_T * O;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
        const int o = r % O, lt = r / O, t = lt % ARMTD_T, l = lt / ARMTD_T;
        const size_t src = size_t(NF) * ARMTD_T_PADDED + (size_t(l) * ARMTD_T_PADDED + t) * O + o;
        if (g) g[r] = g_full[src];
        if (jac)
            for (int v = 0; v < NF; v++) jac[size_t(r) * NF + v] = j_full[src * NF + v];
    }
    if (blockIdx.x == 0 && threadIdx.x < NF) {
        const int i = threadIdx.x;
        double ext[4], grad[4];
        armtd_joint_extremum(A.q0[i], A.qd0[i], A.k_range[i] * k[i], ext, grad);
        for (int q = 0; q < 4; q++) {
            const int row = rows + q * NF + i;
            if (g) g[row] = ext[q];
            if (jac)
                for (int v = 0; v < NF; v++) jac[size_t(row) * NF + v] = (v == i) ? grad[q] : 0.0;
        }
    }
}

}  // namespace armour
