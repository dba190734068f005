// Constraint-side kernels of the hot path (sm_100a, FP64 SIMT, HBM-streaming):
//
//   k_hyperplanes  — K3a: buffered obstacle zonotope -> 36 half-space normals per (link, interval,
//                    obstacle).  Replaces bufferObstaclesKernel + polytope_PH
//                    (reference KPR/CollisionChecking.cu:136-228), which the reference launches 14 times.
//   k_constraints  — K3: slice every link / torque reach set at k (value and d/dk), evaluate the
//                    collision rows and their gradients, append the Bezier joint-limit rows; writes the
//                    whole g(k) and dense Jacobian.  Replaces PZsparse::slice (KPR/PZsparse.cu:404-555),
//                    checkCollisionKernel (KPR/CollisionChecking.cu:230-299) and the row assembly of
//                    armtd_NLP::eval_g / eval_jac_g (KPR/NLPclass.cu:272-396): 1 launch instead of
//                    7 launches + 8-9 blocking PCIe copies per call.
//   k_verdict      — feasibility predicate of armtd_NLP::finalize_solution (KPR/NLPclass.cu:449-537).
//
// One CTA handles TB = 8 consecutive intervals of one problem, so that each output run per link is
// TB*O contiguous doubles (g) or TB*O*7 (Jacobian) and the hyper-plane chunk it streams is contiguous.
// Included by kernels.cu (one translation unit, so all kernels share the __constant__ block).
// Compiled with -fmad=false: coefficient arithmetic is plain round-to-nearest like the host reference.
#pragma once
#include <cuda_runtime.h>

#include "bezier.cuh"
#include "device_constants.cuh"
#include "layout.h"

namespace armour {

// ---------------------------------------------------------------------------------------------------
// K3a
__global__ void __launch_bounds__(256) k_hyperplanes(Batch B) {
    const int tb = blockIdx.x, p = blockIdx.y;
    const int NJ = B.NJ, O = B.O;
    const int per_pair = NJ * TB * O;
    const int items = NCOMB * per_pair;
    const double* obs = B.obstacles + size_t(p) * O * 12;
    const double* LGs = B.link_gens + (size_t(p) * B.T + size_t(tb) * TB) * NJ * 18;
    double* out = B.hp + (size_t(p) * (B.T / TB) + tb) * B.hp_chunk();

    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int i = it / per_pair;
        const int x = it - i * per_pair;  // (l*TB + tt)*O + o
        const int o = x % O;
        const int ltt = x / O;
        const int tt = ltt % TB, l = ltt / TB;
        const double* LG = LGs + size_t(tt * NJ + l) * 18;
        const double* ob = obs + o * 12;
        double G[9][3];
#pragma unroll
        for (int e = 0; e < 3; e++) {
#pragma unroll
            for (int gI = 0; gI < 3; gI++) G[gI][e] = ob[(gI + 1) * 3 + e];
#pragma unroll
            for (int gI = 0; gI < 6; gI++) G[gI + 3][e] = LG[e + gI * 3];
        }
        const int a = c_combA[i], b = c_combB[i];
        double ga[3], gb[3];
#pragma unroll
        for (int e = 0; e < 3; e++) {  // dynamic row select without local-memory indexing
            double va = 0, vb = 0;
#pragma unroll
            for (int gI = 0; gI < 9; gI++) {
                va = (gI == a) ? G[gI][e] : va;
                vb = (gI == b) ? G[gI][e] : vb;
            }
            ga[e] = va;
            gb[e] = vb;
        }
        const double cx = ga[1] * gb[2] - ga[2] * gb[1];
        const double cy = ga[2] * gb[0] - ga[0] * gb[2];
        const double cz = ga[0] * gb[1] - ga[1] * gb[0];
        const double nrm = sqrt(cx * cx + cy * cy + cz * cz);
        double C0 = 0, C1 = 0, C2 = 0;
        if (nrm > 0) {
            C0 = cx / nrm;
            C1 = cy / nrm;
            C2 = cz / nrm;
        }
        const double d = C0 * ob[0] + C1 * ob[1] + C2 * ob[2];
        double delta = 0.0;
#pragma unroll
        for (int gI = 0; gI < 9; gI++) delta += fabs(C0 * G[gI][0] + C1 * G[gI][1] + C2 * G[gI][2]);
        const size_t stride = size_t(NCOMB) * per_pair;
        out[it] = C0;
        out[it + stride] = C1;
        out[it + 2 * stride] = C2;
        out[it + 3 * stride] = d;
        out[it + 4 * stride] = delta;
    }
}

// ---------------------------------------------------------------------------------------------------
// K3
template <int NC>  // NC = 3 (link) or 1 (torque)
__device__ __forceinline__ void slice_one(const uint16_t* __restrict__ keys, const double* __restrict__ coef, int n,
                                          int v, const double (*kp)[4], const double (*dkp)[4], double* acc) {
    // acc[] enters holding the centre (v == 0) or zero (v > 0: derivative w.r.t. k_{v-1})
    for (int mI = 0; mI < n; mI++) {
        const unsigned key = keys[mI];
        double t[NC];
#pragma unroll
        for (int e = 0; e < NC; e++) t[e] = coef[mI * NC + e];
        bool zero = false;
#pragma unroll
        for (int j = 0; j < NF; j++) {
            const int dg = (key >> (2 * j)) & 3;
            double f;
            if (j == v - 1) {
                zero = zero || (dg == 0);
                f = dkp[j][dg];
            } else {
                f = kp[j][dg];
            }
#pragma unroll
            for (int e = 0; e < NC; e++) t[e] *= f;
        }
#pragma unroll
        for (int e = 0; e < NC; e++) acc[e] += zero ? 0.0 : t[e];
    }
}

__global__ void __launch_bounds__(256)
k_constraints(Batch B, const double* __restrict__ kin, double* __restrict__ g, double* __restrict__ jac) {
    const int tb = blockIdx.x, p = blockIdx.y;
    const int NJ = B.NJ, O = B.O, T = B.T;
    const int m = B.m();
    __shared__ double kp[NF][4], dkp[NF][4];
    __shared__ double s_lc[TB][MAXJ][3];
    __shared__ double s_dlc[TB][MAXJ][NF][3];

    if (threadIdx.x < NF) {
        const double k = kin[size_t(p) * NF + threadIdx.x];
        kp[threadIdx.x][0] = 1.0;
        kp[threadIdx.x][1] = k;
        kp[threadIdx.x][2] = k * k;
        kp[threadIdx.x][3] = k * k * k;
        dkp[threadIdx.x][0] = 0.0;
        dkp[threadIdx.x][1] = 1.0;
        dkp[threadIdx.x][2] = 2.0 * k;
        dkp[threadIdx.x][3] = 3.0 * (k * k);
    }
    __syncthreads();

    double* gp = g ? g + size_t(p) * m : nullptr;
    double* jp = jac ? jac + size_t(p) * m * NF : nullptr;

    // phase 1: slices.  item = (tt, s, v): s < NJ link slices, then NF torque slices; v = 0 value, 1..7 d/dk
    const int nsl = NJ + NF;
    for (int it = threadIdx.x; it < TB * nsl * 8; it += blockDim.x) {
        const int v = it & 7;
        const int s = (it >> 3) % nsl;
        const int tt = (it >> 3) / nsl;
        const int t = tb * TB + tt;
        if (s < NJ) {
            const size_t idx = (size_t(p) * T + t) * NJ + s;
            const int n = B.link_n[idx];
            double acc[3] = {0, 0, 0};
            if (v == 0) {
                acc[0] = B.link_c[idx * 3 + 0];
                acc[1] = B.link_c[idx * 3 + 1];
                acc[2] = B.link_c[idx * 3 + 2];
            }
            slice_one<3>(B.link_key + idx * B.capL, B.link_g + idx * B.capL * 3, n, v, kp, dkp, acc);
            if (v == 0) {
                // centre of Interval(c - r, c + r), as getCenter(slice()) does (KPR/NLPclass.cu:313)
                const double* LG = B.link_gens + idx * 18;
#pragma unroll
                for (int e = 0; e < 3; e++) {
                    const double r = LG[e + (3 + e) * 3];
                    const double c = ((acc[e] - r) + (acc[e] + r)) * 0.5;
                    s_lc[tt][s][e] = c;
                    B.link_sliced[idx * 3 + e] = c;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 3; e++) s_dlc[tt][s][v - 1][e] = acc[e];
            }
        } else {
            const int j = s - NJ;
            const size_t idx = (size_t(p) * T + t) * NF + j;
            const int n = B.u_n[idx];
            double acc[1] = {v == 0 ? B.u_c[idx] : 0.0};
            slice_one<1>(B.u_key + idx * B.capU, B.u_g + idx * B.capU, n, v, kp, dkp, acc);
            if (v == 0) {
                const double r = B.u_r[idx];
                if (gp) gp[t * NF + j] = ((acc[0] - r) + (acc[0] + r)) * 0.5;
            } else if (jp) {
                jp[size_t(t * NF + j) * NF + (v - 1)] = acc[0];
            }
        }
    }
    __syncthreads();

    // phase 2: collision rows.  item x = (l*TB + tt)*O + o
    const int per_pair = NJ * TB * O;
    const double* hp = B.hp + (size_t(p) * (T / TB) + tb) * B.hp_chunk();
    const size_t cstride = size_t(NCOMB) * per_pair;
    for (int x = threadIdx.x; x < per_pair; x += blockDim.x) {
        const int o = x % O;
        const int ltt = x / O;
        const int tt = ltt % TB, l = ltt / TB;
        const double c0 = s_lc[tt][l][0], c1 = s_lc[tt][l][1], c2 = s_lc[tt][l][2];
        double max_elt = -100000000;
        double A0 = 0, A1 = 0, A2 = 0;  // winning normal, sign folded in
#pragma unroll 4
        for (int i = 0; i < NCOMB; i++) {
            const double* h = hp + size_t(i) * per_pair + x;
            const double a0 = h[0], a1 = h[cstride], a2 = h[2 * cstride];
            const double d = h[3 * cstride], dl = h[4 * cstride];
            double pos = -100000000, neg = -100000000;
            if (sqrt(a0 * a0 + a1 * a1 + a2 * a2) > 0) {
                const double dot = a0 * c0 + a1 * c1 + a2 * c2;
                pos = dot - (d + dl);
                neg = -dot - (-d + dl);
            }
            if (pos > max_elt) {  // strict '>' and pos-before-neg: KPR/CollisionChecking.cu:264-276
                max_elt = pos;
                A0 = -a0; A1 = -a1; A2 = -a2;
            }
            if (neg > max_elt) {
                max_elt = neg;
                A0 = a0; A1 = a1; A2 = a2;
            }
        }
        const int t = tb * TB + tt;
        const size_t row = size_t(NF) * T + (size_t(l) * T + t) * O + o;
        if (gp) gp[row] = -max_elt;
        if (jp) {
#pragma unroll
            for (int v = 0; v < NF; v++) {
                const double* dk = s_dlc[tt][l][v];
                // -(A.dk) for a 'pos' winner, +(A.dk) for 'neg' (:286-295); the sign was folded into A
                jp[row * NF + v] = A0 * dk[0] + A1 * dk[1] + A2 * dk[2];
            }
        }
    }

    // Bezier joint-limit rows (KPR/Trajectory.cu:256-540), once per problem
    if (tb == 0 && threadIdx.x < NF) {
        const int i = threadIdx.x;
        const double D = c_robot.duration;
        const double q0 = B.q0[size_t(p) * NF + i];
        const double a = B.qd0[size_t(p) * NF + i] * D;
        const double b = B.qdd0[size_t(p) * NF + i] * D * D;
        const double kn = kin[size_t(p) * NF + i];
        const int off = NF * T + NJ * T * O;
        for (int vel = 0; vel < 2; vel++) {
            double mn, mx, dmn, dmx;
            bez_extrema(vel == 1, q0, a, b, c_robot.k_range[i], D, kn, &mn, &mx, &dmn, &dmx);
            const int r0 = off + vel * 2 * NF + i;
            if (gp) {
                gp[r0] = mn;
                gp[r0 + NF] = mx;
            }
            if (jp) {
                for (int j = 0; j < NF; j++) {
                    jp[size_t(r0) * NF + j] = (j == i) ? dmn : 0.0;
                    jp[size_t(r0 + NF) * NF + j] = (j == i) ? dmx : 0.0;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// verdict: first violated row in the reference's check order (= ascending row index), one CTA per problem
__global__ void __launch_bounds__(256)
k_verdict(Batch B, const double* __restrict__ g, int* __restrict__ feasible, int* __restrict__ first) {
    const int p = blockIdx.x;
    const int T = B.T, NJ = B.NJ, O = B.O, m = B.m();
    const double* gp = g + size_t(p) * m;
    const double* tr = B.torque_radius + size_t(p) * NF * T;
    int best = m;
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        const double v = gp[r];
        bool bad;
        if (r < NF * T) {
            const int t = r / NF, j = r % NF;
            const double rad = tr[j * T + t];
            bad = v < -c_robot.torque_limits[j] + rad - c_robot.torque_violation_threshold ||
                  v > c_robot.torque_limits[j] - rad + c_robot.torque_violation_threshold;
        } else if (r < NF * T + NJ * T * O) {
            bad = v > c_robot.collision_violation_threshold;
        } else {
            const int q = r - (NF * T + NJ * T * O);
            const int j = q % NF;
            if (q < 2 * NF)
                bad = v < c_robot.state_limits_lb[j] + c_robot.qe || v > c_robot.state_limits_ub[j] - c_robot.qe;
            else
                bad = v < -c_robot.speed_limits[j] + c_robot.qde || v > c_robot.speed_limits[j] - c_robot.qde;
        }
        if (bad && r < best) best = r;
    }
    __shared__ int s_best;
    if (threadIdx.x == 0) s_best = m;
    __syncthreads();
    atomicMin(&s_best, best);
    __syncthreads();
    if (threadIdx.x == 0) {
        feasible[p] = (s_best == m) ? 1 : 0;
        first[p] = (s_best == m) ? -1 : s_best;
    }
}

// ---------------------------------------------------------------------------------------------------
// host launchers (called from capi.cu)
cudaError_t launch_hyperplanes(const Batch& B, cudaStream_t st) {
    if (B.O == 0 || B.nprob == 0) return cudaSuccess;
    dim3 grid(B.T / TB, B.nprob);
    k_hyperplanes<<<grid, 256, 0, st>>>(B);
    return cudaGetLastError();
}
cudaError_t launch_constraints(const Batch& B, const double* d_k, double* d_g, double* d_jac, cudaStream_t st) {
    if (B.nprob == 0) return cudaSuccess;
    dim3 grid(B.T / TB, B.nprob);
    k_constraints<<<grid, 256, 0, st>>>(B, d_k, d_g, d_jac);
    return cudaGetLastError();
}
cudaError_t launch_verdict(const Batch& B, const double* d_g, int* d_feasible, int* d_first, cudaStream_t st) {
    if (B.nprob == 0) return cudaSuccess;
    k_verdict<<<B.nprob, 256, 0, st>>>(B, d_g, d_feasible, d_first);
    return cudaGetLastError();
}

}  // namespace armour
