// Constraint-side kernels of the hot path (sm_100a, FP64 SIMT, HBM-streaming):
//
//   k_hyperplanes  — K3a: buffered obstacle zonotope -> the half-spaces that can attain the maximum of a
//                    (link, interval, obstacle) row for some k in the box.  Replaces bufferObstaclesKernel +
//                    polytope_PH (reference KPR/CollisionChecking.cu:136-228, 14 launches).
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
// Collision half-spaces of one (link, interval, obstacle) row.
//
// The buffered obstacle zonotope has 9 generators: 3 of the obstacle, the <= 3 pure link-generator
// columns and diag(radius) of the link reach set (KPR/CollisionChecking.cu:136-167).  Each of the
// C(9,2) = 36 generator pairs gives a unit normal C = g_a x g_b / |g_a x g_b| (zero if degenerate),
// an offset d = C . c_obs and a half-width delta = sum_j |C . g_j| (polytope_PH, :169-228); the
// constraint of the row is h(k) = -max_i max(C_i.p(k) - (d_i + delta_i), -C_i.p(k) - (-d_i + delta_i))
// with strict '>' in the scan order pos_0, neg_0, pos_1, ... (checkCollisionKernel, :230-299).
// for_each_plane() enumerates the pairs in the reference order (0,1),(0,2)...(7,8) with every index a
// compile-time constant, so the 27 generator components stay in registers.
template <class F>
__device__ __forceinline__ void for_each_plane(const double (&G)[9][3], const double (&oc)[3], F f) {
    int i = 0;
#pragma unroll
    for (int a = 0; a < 8; a++) {
#pragma unroll
        for (int b = a + 1; b < 9; b++) {
            const double cx = G[a][1] * G[b][2] - G[a][2] * G[b][1];
            const double cy = G[a][2] * G[b][0] - G[a][0] * G[b][2];
            const double cz = G[a][0] * G[b][1] - G[a][1] * G[b][0];
            const double nrm = sqrt(cx * cx + cy * cy + cz * cz);
            double C0 = 0, C1 = 0, C2 = 0;
            if (nrm > 0) {
                C0 = cx / nrm;
                C1 = cy / nrm;
                C2 = cz / nrm;
            }
            const double d = C0 * oc[0] + C1 * oc[1] + C2 * oc[2];
            double delta = 0.0;
#pragma unroll
            for (int gI = 0; gI < 9; gI++) delta += fabs(C0 * G[gI][0] + C1 * G[gI][1] + C2 * G[gI][2]);
            const bool nz = sqrt(C0 * C0 + C1 * C1 + C2 * C2) > 0;  // the test of checkCollisionKernel (:259)
            f(i, C0, C1, C2, d, delta, nz);
            i++;
        }
    }
}

__device__ __forceinline__ void load_row_generators(const double* __restrict__ ob, const double* __restrict__ LG,
                                                    double (&G)[9][3], double (&oc)[3]) {
#pragma unroll
    for (int e = 0; e < 3; e++) {
        oc[e] = ob[e];
#pragma unroll
        for (int gI = 0; gI < 3; gI++) G[gI][e] = ob[(gI + 1) * 3 + e];
#pragma unroll
        for (int gI = 0; gI < 6; gI++) G[gI + 3][e] = LG[e + gI * 3];
    }
}

// ---------------------------------------------------------------------------------------------------
// K3a: half-space candidate lists.  Replaces bufferObstaclesKernel + polytope_PH (14 launches in the
// reference).  The reference stores all 72 half-spaces of every row (12.9 MB per problem at 10
// obstacles) and scans them at every constraint evaluation.  Here the scan is bounded at build time:
// the link centre p(k) is a polynomial in k with known coefficients, so each half-space value
//     v_i(k) = s_i C_i . p(k) - b_i     lies in   [vc_i - rho_i, vc_i + rho_i]   for all |k_j| <= K_DOMAIN,
// rho_i = sum_m |C_i . g_m|.  A half-space whose upper bound is below the best lower bound can never
// be the maximum; neither can one that repeats the normal of an earlier candidate with a larger
// offset.  Only the survivors are stored, in scan order, as 32-byte records (s_i C_i, b_i), so that the
// evaluation kernel returns bit-identical values and gradients while reading ~100 B per row instead
// of 1440 B.  Rows with more than HP_CAP survivors are flagged and evaluated from the generators.
__global__ void __launch_bounds__(128) k_hyperplanes(Batch B) {
    const int tb = blockIdx.x, p = blockIdx.y;
    const int NJ = B.NJ, O = B.O, T = B.T;
    const int per_pair = NJ * TB * O;
    const double* obs = B.obstacles + size_t(p) * O * 12;
    const size_t chunk = size_t(p) * (T / TB) + tb;
    double* cand = B.hp_cand + chunk * B.hp_chunk();
    unsigned char* cnt = B.hp_cnt + chunk * per_pair;

    for (int x = threadIdx.x; x < per_pair; x += blockDim.x) {
        const int o = x % O;
        const int ltt = x / O;
        const int tt = ltt % TB, l = ltt / TB;
        const size_t idx = (size_t(p) * T + size_t(tb) * TB + tt) * NJ + l;
        double G[9][3], oc[3];
        load_row_generators(obs + o * 12, B.link_gens + idx * 18, G, oc);
        const int n = B.link_n[idx];
        const double* __restrict__ lg = B.link_g + idx * B.capL * 3;
        const double c0 = B.link_c[idx * 3 + 0], c1 = B.link_c[idx * 3 + 1], c2 = B.link_c[idx * 3 + 2];

        auto bounds = [&](double C0, double C1, double C2, double d, double delta, double& vpos, double& vneg,
                          double& rho) {
            const double dot = C0 * c0 + C1 * c1 + C2 * c2;
            vpos = dot - (d + delta);
            vneg = -dot - (-d + delta);
            double r = 0.0;
            for (int mI = 0; mI < n; mI++)
                r += fabs(C0 * lg[mI * 3 + 0] + C1 * lg[mI * 3 + 1] + C2 * lg[mI * 3 + 2]);
            // |k_j| <= K_DOMAIN, total degree <= 21; plus evaluation round-off (values are O(1))
            rho = r * HP_RHO_SCALE + (1e-10 + 1e-12 * (fabs(dot) + fabs(d) + delta));
        };

        // pass 1: the best guaranteed lower bound
        double lo_max = -100000000;
        for_each_plane(G, oc, [&](int, double C0, double C1, double C2, double d, double delta, bool nz) {
            if (!nz) return;
            double vpos, vneg, rho;
            bounds(C0, C1, C2, d, delta, vpos, vneg, rho);
            lo_max = fmax(lo_max, fmax(vpos, vneg) - rho);
        });
        // pass 2: emit the survivors in scan order
        int count = 0;
        bool overflow = false;
        double* row = cand + size_t(x) * 4;
        const size_t cstride = size_t(per_pair) * 4;
        auto emit = [&](double A0, double A1, double A2, double b) {
            for (int q = 0; q < count; q++) {  // an earlier candidate with the same normal and b_q <= b dominates
                const double* e = row + q * cstride;
                if (e[0] == A0 && e[1] == A1 && e[2] == A2 && e[3] <= b) return;
            }
            if (count == HP_CAP) {
                overflow = true;
                return;
            }
            double* e = row + count * cstride;
            e[0] = A0;
            e[1] = A1;
            e[2] = A2;
            e[3] = b;
            count++;
        };
        for_each_plane(G, oc, [&](int, double C0, double C1, double C2, double d, double delta, bool nz) {
            if (!nz || overflow) return;
            double vpos, vneg, rho;
            bounds(C0, C1, C2, d, delta, vpos, vneg, rho);
            if (vpos + rho >= lo_max) emit(C0, C1, C2, d + delta);
            if (vneg + rho >= lo_max) emit(-C0, -C1, -C2, -d + delta);
        });
        cnt[x] = overflow ? (unsigned char)HP_OVERFLOW : (unsigned char)count;
    }
}

// ---------------------------------------------------------------------------------------------------
// K3
// Slow path of one collision row: scan all 72 half-spaces computed from the generators
// (checkCollisionKernel, KPR/CollisionChecking.cu:230-299).  Kept out of line so that the streaming
// path of k_constraints keeps its register budget.
__device__ __noinline__ void row_from_generators(const double* __restrict__ ob, const double* __restrict__ LG,
                                                 double c0, double c1, double c2, double* max_out, double* A0o,
                                                 double* A1o, double* A2o) {
    double G[9][3], oc[3];
    load_row_generators(ob, LG, G, oc);
    double max_elt = -100000000, A0 = 0, A1 = 0, A2 = 0;
    for_each_plane(G, oc, [&](int, double a0, double a1, double a2, double d, double dl, bool nz) {
        if (!nz) return;
        const double dot = a0 * c0 + a1 * c1 + a2 * c2;
        const double pos = dot - (d + dl);
        const double neg = -dot - (-d + dl);
        if (pos > max_elt) {
            max_elt = pos;
            A0 = -a0; A1 = -a1; A2 = -a2;
        }
        if (neg > max_elt) {
            max_elt = neg;
            A0 = a0; A1 = a1; A2 = a2;
        }
    });
    *max_out = max_elt;
    *A0o = A0;
    *A1o = A1;
    *A2o = A2;
}

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

__global__ void __launch_bounds__(256, 4)
k_constraints(Batch B, const double* __restrict__ kin, double* __restrict__ g, double* __restrict__ jac) {
    const int tb = blockIdx.x, p = blockIdx.y;
    const int NJ = B.NJ, O = B.O, T = B.T;
    const int m = B.m();
    __shared__ double kp[NF][4], dkp[NF][4];
    __shared__ double s_lc[TB][MAXJ][3];
    __shared__ double s_dlc[TB][MAXJ][NF][3];
    __shared__ int s_in_domain;

    if (threadIdx.x == 0) {
        bool in = true;
        for (int j = 0; j < NF; j++) in = in && (fabs(kin[size_t(p) * NF + j]) <= K_DOMAIN);
        s_in_domain = in ? 1 : 0;
    }
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
    const size_t chunk = size_t(p) * (T / TB) + tb;
    const double* cand = B.hp_cand + chunk * B.hp_chunk();
    const unsigned char* cnt = B.hp_cnt + chunk * per_pair;
    const size_t cstride = size_t(per_pair) * 4;
    for (int x = threadIdx.x; x < per_pair; x += blockDim.x) {
        const int o = x % O;
        const int ltt = x / O;
        const int tt = ltt % TB, l = ltt / TB;
        const double c0 = s_lc[tt][l][0], c1 = s_lc[tt][l][1], c2 = s_lc[tt][l][2];
        double max_elt = -100000000;
        double A0 = 0, A1 = 0, A2 = 0;  // minus the winning signed normal
        const int n = cnt[x];
        if (n != HP_OVERFLOW && s_in_domain) {
            const double2* row = reinterpret_cast<const double2*>(cand + size_t(x) * 4);
            for (int q = 0; q < n; q++) {
                const double2 u = __ldg(row + q * (cstride / 2));
                const double2 w = __ldg(row + q * (cstride / 2) + 1);
                const double v = (u.x * c0 + u.y * c1 + w.x * c2) - w.y;
                if (v > max_elt) {  // strict '>' in scan order: KPR/CollisionChecking.cu:264-276
                    max_elt = v;
                    A0 = -u.x; A1 = -u.y; A2 = -w.x;
                }
            }
        } else {
            // all 72 half-spaces from the generators: rows with more than HP_CAP candidates, or k outside
            // the box the candidate lists were built for
            const size_t idx = (size_t(p) * T + tb * TB + tt) * NJ + l;
            row_from_generators(B.obstacles + (size_t(p) * O + o) * 12, B.link_gens + idx * 18, c0, c1, c2, &max_elt,
                                &A0, &A1, &A2);
        }
        const int t = tb * TB + tt;
        const size_t row_i = size_t(NF) * T + (size_t(l) * T + t) * O + o;
        if (gp) gp[row_i] = -max_elt;
        if (jp) {
#pragma unroll
            for (int v = 0; v < NF; v++) {
                const double* dk = s_dlc[tt][l][v];
                // -(C.dk) for a 'pos' winner, +(C.dk) for 'neg' (:286-295); the sign is folded into A
                jp[row_i * NF + v] = A0 * dk[0] + A1 * dk[1] + A2 * dk[2];
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
    k_hyperplanes<<<grid, 128, 0, st>>>(B);
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
