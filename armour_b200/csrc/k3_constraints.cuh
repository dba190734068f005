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

// tuning knobs of k_constraints (see the sweep quoted at the kernel)
#ifndef K3_CH
#define K3_CH 2   // monomials of a table fetched together by a slicing thread
#endif
#ifndef K3_CQ
#define K3_CQ 2   // candidate records in flight per collision row (2.05 stored per row on average)
#endif
#ifndef K3_MINB
#define K3_MINB 4 // resident CTAs per SM the register budget is sized for (48 registers per thread)
#endif

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
// One thread per (row, generator pair): HP_ROWS rows per CTA.  Step 1 computes the pair's two half-spaces and
// their value bounds, step 2 the row's best guaranteed lower bound, step 3 decides per candidate (bound
// filter + "an earlier candidate with the same normal and a smaller offset dominates"; dominance is
// transitive, so testing against all earlier candidates that pass the filter equals the sequential rule),
// step 4 writes the survivors at their scan-order positions.  Bit-identical lists to a sequential scan.
constexpr int HP_ROWS = 8;
constexpr int HP_STAGE = 24;  // link monomials staged in shared memory per row (longer tables are read from global)
constexpr int HP_THREADS = HP_ROWS * NCOMB;  // 288
__global__ void __launch_bounds__(HP_THREADS) k_hyperplanes(Batch B) {
    const int p = blockIdx.y;
    const int NJ = B.NJ, O = B.O, T = B.T;
    const int per_pair = NJ * TB * O;          // rows of one (problem, TB intervals) chunk
    const int rows_total = per_pair * (T / TB);
    const int lr = threadIdx.x / NCOMB, pi = threadIdx.x % NCOMB;  // local row, pair index
    const int r = blockIdx.x * HP_ROWS + lr;   // row of the problem, chunk-major
    __shared__ double s_G[HP_ROWS][9][3], s_oc[HP_ROWS][3], s_c[HP_ROWS][3];
    __shared__ double s_A[HP_ROWS][NCOMB][3], s_b[HP_ROWS][2 * NCOMB], s_up[HP_ROWS][2 * NCOMB], s_lo[HP_ROWS][NCOMB];
    __shared__ double s_lomax[HP_ROWS];
    __shared__ double s_vc[HP_ROWS][2 * NCOMB];           // value of each signed half-space at the centre of the k box
    __shared__ double s_gm[HP_ROWS][HP_STAGE][3];        // monomial coefficients of the row's link reach set
    __shared__ int s_best[HP_ROWS];                        // signed half-space with the best guaranteed lower bound
    __shared__ unsigned char s_flag[HP_ROWS][2 * NCOMB], s_list[HP_ROWS][2 * NCOMB];
    __shared__ int s_nlist[HP_ROWS];
    const bool live = r < rows_total;
    int tb = 0, x = 0;
    size_t idx = 0;
    if (live) {
        tb = r / per_pair;
        x = r - tb * per_pair;
        const int o = x % O, ltt = x / O, tt = ltt % TB, l = ltt / TB;
        idx = (size_t(p) * T + size_t(tb) * TB + tt) * NJ + l;
        if (pi < 27) {  // 9 generators x 3 components: 3 of the obstacle, then the 3x6 link matrix (column-major)
            const int gI = pi / 3, e = pi % 3;
            s_G[lr][gI][e] = (gI < 3) ? B.obstacles[(size_t(p) * O + o) * 12 + (gI + 1) * 3 + e] : B.link_gens[idx * 18 + e + (gI - 3) * 3];
        } else if (pi < 30) {
            s_oc[lr][pi - 27] = B.obstacles[(size_t(p) * O + o) * 12 + (pi - 27)];
        } else if (pi < 33) {
            s_c[lr][pi - 30] = B.link_c[idx * 3 + (pi - 30)];
        } else if (pi == 33) {
            s_nlist[lr] = 0;
        }
        const int n = B.link_n[idx];
        const int ns = n < HP_STAGE ? n : HP_STAGE;
        for (int q = pi; q < ns * 3; q += NCOMB) s_gm[lr][q / 3][q % 3] = B.link_g[idx * B.capL * 3 + q];
    }
    __syncthreads();
    bool nz = false;
    if (live) {  // step 1: same expressions as for_each_plane / the bounds of the sequential version
        const int a = c_combA[pi], b = c_combB[pi];
        const double(*G)[3] = s_G[lr];
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
        const double d = C0 * s_oc[lr][0] + C1 * s_oc[lr][1] + C2 * s_oc[lr][2];
        double delta = 0.0;
#pragma unroll 1
        for (int gI = 0; gI < 9; gI++) delta += fabs(C0 * G[gI][0] + C1 * G[gI][1] + C2 * G[gI][2]);
        nz = sqrt(C0 * C0 + C1 * C1 + C2 * C2) > 0;  // the test of checkCollisionKernel (:259)
        const double dot = C0 * s_c[lr][0] + C1 * s_c[lr][1] + C2 * s_c[lr][2];
        const double vpos = dot - (d + delta);
        const double vneg = -dot - (-d + delta);
        const int n = B.link_n[idx];
        const int ns = n < HP_STAGE ? n : HP_STAGE;
        const double* __restrict__ lg = B.link_g + idx * B.capL * 3;
        double rr = 0.0;
#pragma unroll 1
        for (int mI = 0; mI < ns; mI++) rr += fabs(C0 * s_gm[lr][mI][0] + C1 * s_gm[lr][mI][1] + C2 * s_gm[lr][mI][2]);
#pragma unroll 1
        for (int mI = ns; mI < n; mI++) rr += fabs(C0 * lg[mI * 3 + 0] + C1 * lg[mI * 3 + 1] + C2 * lg[mI * 3 + 2]);
        // |k_j| <= K_DOMAIN, total degree <= 21; plus evaluation round-off (values are O(1))
        const double rho = rr * HP_RHO_SCALE + (1e-10 + 1e-12 * (fabs(dot) + fabs(d) + delta));
        s_A[lr][pi][0] = C0;
        s_A[lr][pi][1] = C1;
        s_A[lr][pi][2] = C2;
        s_b[lr][2 * pi] = d + delta;
        s_b[lr][2 * pi + 1] = -d + delta;
        s_up[lr][2 * pi] = nz ? vpos + rho : -1e300;
        s_up[lr][2 * pi + 1] = nz ? vneg + rho : -1e300;
        s_lo[lr][pi] = nz ? fmax(vpos, vneg) - rho : -1e300;
        s_vc[lr][2 * pi] = vpos;
        s_vc[lr][2 * pi + 1] = vneg;
    }
    __syncthreads();
    if (live && pi == 0) {  // step 2: the best guaranteed lower bound of the row, and who holds it
        double lo_max = -100000000;
        int best = -1;
#pragma unroll 1
        for (int i = 0; i < NCOMB; i++) {
            if (s_lo[lr][i] > lo_max) {
                lo_max = s_lo[lr][i];
                best = 2 * i + (s_vc[lr][2 * i + 1] > s_vc[lr][2 * i] ? 1 : 0);
            }
        }
        s_lomax[lr] = lo_max;
        s_best[lr] = best;
    }
    __syncthreads();
    // step 2b: pairwise test against that half-space j.  v_j(k) - v_i(k) = (vc_j - vc_i) + (A_j - A_i) . sum_m g_m mono_m(k)
    // >= (vc_j - vc_i) - sum_m |(A_j - A_i) . g_m| for every k of the box: the variation the two half-spaces share
    // cancels, so this removes the half-spaces that are nearly parallel to j but further out, which the separate
    // bounds of step 1 cannot.  i is dropped only when the difference is positive by a margin (never on a tie), so
    // the maximum over the survivors and its first attaining index are those of the full scan.
    bool drop[2] = {false, false};
    if (live && nz && s_best[lr] >= 0) {
        const int jb = s_best[lr];
        const double sj = (jb & 1) ? -1.0 : 1.0;
        const double J0 = sj * s_A[lr][jb >> 1][0], J1 = sj * s_A[lr][jb >> 1][1], J2 = sj * s_A[lr][jb >> 1][2];
        const double vj = s_vc[lr][jb];
        const int n = B.link_n[idx];
        const int ns = n < HP_STAGE ? n : HP_STAGE;
        const double* __restrict__ lg = B.link_g + idx * B.capL * 3;
#pragma unroll 1
        for (int sgn = 0; sgn < 2; sgn++) {
            const int sI = 2 * pi + sgn;
            if (sI == jb || !(s_up[lr][sI] >= s_lomax[lr])) continue;
            const double sg = sgn ? -1.0 : 1.0;
            const double D0 = J0 - sg * s_A[lr][pi][0], D1 = J1 - sg * s_A[lr][pi][1], D2 = J2 - sg * s_A[lr][pi][2];
            double rd = 0.0;
#pragma unroll 1
            for (int mI = 0; mI < ns; mI++) rd += fabs(D0 * s_gm[lr][mI][0] + D1 * s_gm[lr][mI][1] + D2 * s_gm[lr][mI][2]);
#pragma unroll 1
            for (int mI = ns; mI < n; mI++) rd += fabs(D0 * lg[mI * 3 + 0] + D1 * lg[mI * 3 + 1] + D2 * lg[mI * 3 + 2]);
            const double vi = s_vc[lr][sI];
            const double margin = rd * HP_RHO_SCALE + (1e-10 + 1e-12 * (fabs(vi) + fabs(vj) + fabs(s_b[lr][sI]) + fabs(s_b[lr][jb])));
            drop[sgn] = (vj - vi) > margin;
        }
    }
    __syncthreads();
    if (live) {
        if (drop[0]) s_up[lr][2 * pi] = -1e300;
        if (drop[1]) s_up[lr][2 * pi + 1] = -1e300;
    }
    __syncthreads();
    // step 3: the candidates that pass the filters enter a short list per row (a handful of the 72; arrival order,
    // the tests below do not depend on it)
    if (live) {
        const double lo_max = s_lomax[lr];
#pragma unroll 1
        for (int sgn = 0; sgn < 2; sgn++) {
            const int sI = 2 * pi + sgn;  // position in the scan order pos_0, neg_0, pos_1, ...
            if (nz && s_up[lr][sI] >= lo_max) s_list[lr][atomicAdd(&s_nlist[lr], 1)] = (unsigned char)sI;
        }
    }
    __syncthreads();
    // step 3b: a listed candidate is dropped when a listed candidate EARLIER in the scan order has the same normal and
    // a smaller or equal offset (it can never win the strict '>' scan).  Dominance is transitive, so testing against
    // all earlier listed candidates equals the sequential rule.
    bool keep[2] = {false, false};
    if (live) {
        const double lo_max = s_lomax[lr];
        const int nl = s_nlist[lr];
#pragma unroll 1
        for (int sgn = 0; sgn < 2; sgn++) {
            const int sI = 2 * pi + sgn;
            bool k = nz && s_up[lr][sI] >= lo_max;
            if (k) {
                const double sg = sgn ? -1.0 : 1.0;
                const double A0 = sg * s_A[lr][pi][0], A1 = sg * s_A[lr][pi][1], A2 = sg * s_A[lr][pi][2], bb = s_b[lr][sI];
#pragma unroll 1
                for (int c = 0; c < nl && k; c++) {
                    const int q = s_list[lr][c];
                    if (q >= sI) continue;
                    const double sq = (q & 1) ? -1.0 : 1.0;
                    const int qi = q >> 1;
                    if (sq * s_A[lr][qi][0] == A0 && sq * s_A[lr][qi][1] == A1 && sq * s_A[lr][qi][2] == A2 && s_b[lr][q] <= bb)
                        k = false;
                }
            }
            keep[sgn] = k;
            s_flag[lr][sI] = k ? 1 : 0;
        }
    }
    __syncthreads();
    // step 4: the row's list is packed behind the lists already placed in its chunk (one atomic per row: the order
    // of the rows inside a chunk is arrival order and differs run to run, the content of every list does not), so
    // that an evaluation fetches exactly the records in use with one bulk copy.  Survivors keep their scan order.
    __shared__ int s_off[HP_ROWS];
    int count = 0, before[2] = {0, 0};
    if (live && (keep[0] || keep[1] || pi == 0)) {
        const int nl = s_nlist[lr];
#pragma unroll 1
        for (int c = 0; c < nl; c++) {
            const int q = s_list[lr][c];
            if (s_flag[lr][q]) {
                count++;
                before[0] += (q < 2 * pi);
                before[1] += (q < 2 * pi + 1);
            }
        }
    }
    const size_t chunk = size_t(p) * (T / TB) + tb;
    if (live && pi == 0) {
        int off = -1;
        if (count <= HP_CAP) {
            off = atomicAdd(B.hp_total_of(p, tb), count);
            if (size_t(off) + count > B.hp_chunk_records()) off = -1;  // chunk full: the row goes to the slow path
        }
        s_off[lr] = off;
        if (off < 0) atomicAdd(B.hp_slow_of(p), 1);
        B.hp_meta[chunk * per_pair + x] = off < 0 ? unsigned(HP_OVERFLOW) : ((unsigned(off) << 8) | unsigned(count));
    }
    __syncthreads();
    if (live && s_off[lr] >= 0) {
        double* row = B.hp_cand + (chunk * B.hp_chunk_records() + size_t(s_off[lr])) * 4;
#pragma unroll 1
        for (int sgn = 0; sgn < 2; sgn++) {
            if (keep[sgn]) {
                const double sg = sgn ? -1.0 : 1.0;
                double* e = row + size_t(before[sgn]) * 4;
                e[0] = sg * s_A[lr][pi][0];
                e[1] = sg * s_A[lr][pi][1];
                e[2] = sg * s_A[lr][pi][2];
                e[3] = s_b[lr][2 * pi + sgn];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K3
// Slow path of one collision row: scan all 72 half-spaces computed from the generators
// (checkCollisionKernel, KPR/CollisionChecking.cu:230-299).  Only k_constraints_slow calls it.
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

// ---- TMA bulk copies (cp.async.bulk global -> shared, completion on an mbarrier) ---------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ unsigned r16(unsigned bytes) { return (bytes + 15u) & ~15u; }

// k_constraints: persistent CTAs (two per SM); each walks a contiguous range of chunks = (problem, TB intervals) as a
// software pipeline fed by TMA bulk copies, so that neither the DRAM latency of a chunk nor the uneven length of its
// pieces leaves warps idle.
//
// Loads of chunk c (all cp.async.bulk into shared memory, completion on mbarriers, buffers by parity of c):
//   G1(c)   fixed sizes: monomial counts, centres and radii of its TB*(NJ+NF) reach-set tables;
//   TAB(c)  every table, exactly its n monomials (16-bit keys, coefficients) — sizes known from G1(c);
//   ROW(c)  the (first record, count) words of its NJ*TB*O collision rows and the candidate half-space records in use of
//           the chunk: one contiguous run.
// Iteration j of a CTA works on TWO chunks at once — the slices of chunk j and the collision rows of chunk j-1 — as one
// pool of items handed to its warps by ticket: 7 torque + NJ link slice items (a warp slices the tables of four
// consecutive intervals of one joint / link, eight lanes per table, lane v owns output v = the value or d/dk_{v-1}),
// NJ*TB*O/32 row tiles (one lane per collision row), the torque rows of chunk j-1, and the placement + issue of TAB(j+1).
// One __syncthreads per iteration.  G1(j+1) and ROW(j) are issued at the top of iteration j.
// The power product of a monomial comes from two tables in shared memory, k0..k3 (256 entries) and k4..k6 (64), whose
// columns hold the plain product and its derivatives: term_v = coeff * A[key & 255][colA(v)] * B[key >> 8][colB(v)].
// They depend on k only, i.e. on the problem: there is one pair per chunk parity, rebuilt when the chunk range crosses
// into the next problem.  (The reference applies the factors one variable after the other, KPR/PZsparse.cu:404-555:
// same value up to a few ulp.)
#ifndef K3_THREADS_N
#define K3_THREADS_N 384
#endif
constexpr int K3_THREADS = K3_THREADS_N;
constexpr int K3_WARPS = K3_THREADS / 32;
constexpr int K3_TABLE_ARENA = 12288;      // bytes of staged tables per stage
constexpr int K3_NTAB = TB * (MAXJ + NF);  // tables of a chunk (upper bound)
static_assert(TB % 4 == 0, "a warp slices four intervals of a link / joint together");
static_assert(K3_NTAB <= 64, "one warp issues the table copies, two tables per lane");
static_assert(K3_THREADS >= 256, "the power-product tables are built by 256 threads");

struct K3Stage {  // slice side of a chunk
    int nl[TB * MAXJ], nu[TB * NF + 4];
    double cen_l[TB * MAXJ * 3], rad_l[TB * MAXJ * 3], cen_u[TB * NF + 4], rad_u[TB * NF + 4];
    int toff[K3_NTAB];  // byte offset of a staged table in the arena, -1: read it from global memory
    int pad[4];
    __align__(16) unsigned char arena[K3_TABLE_ARENA];
};
struct K3Rows {  // row side of a chunk: what its slices leave for the collision rows
    double lc[TB][MAXJ][4];       // sliced link centres
    double dlc[TB][MAXJ][NF][4];  // and their d/dk
    double tg[TB * NF + 4];       // torque rows of g
    double tj[TB * NF * NF + 4];  // torque rows of the Jacobian
    int staged, in_domain, pad[2];
};
struct K3Smem {  // fixed part of the dynamic shared memory; 2 x row words and 2 x candidate arena follow
    unsigned long long bar_g1[2], bar_tab[2], bar_row[2];
    double pwA[256][5];  // k0..k3: plain product, d/dk0 .. d/dk3
    double pwB[64][4];   // k4..k6: plain product, d/dk4 .. d/dk6
    double k[8];
    double zero_d[4];    // operands of an empty table
    unsigned short zero_k[8];
    int ticket, pad[3];
    K3Stage st[2];
    K3Rows rw[2];
};

// one table sliced by an 8-lane group, operands in shared memory.  n >= 0 own monomials, nmax = longest of the warp's
// four tables (lanes past their own n re-read their last monomial with a zero factor: no branch, no garbage).  Two
// accumulator sets: consecutive monomials do not wait for each other's multiply-add.
template <bool TORQUE>
__device__ __forceinline__ void slice_group(const unsigned short* __restrict__ kp, const double* __restrict__ cp, int n,
                                            int nmax, const double* __restrict__ pa, const double* __restrict__ pb,
                                            double& a0, double& a1, double& a2) {
    const int nlast = n > 0 ? n - 1 : 0;
    double b0 = 0.0, b1 = 0.0, b2 = 0.0;
    int mI = 0;
    for (; mI + 1 < nmax; mI += 2) {
        const int m0 = mI < nlast ? mI : nlast, m1 = mI + 1 < nlast ? mI + 1 : nlast;
        const unsigned key0 = kp[m0], key1 = kp[m1];
        double f0 = pa[(key0 & 255u) * 5] * pb[(key0 >> 8) * 4];
        double f1 = pa[(key1 & 255u) * 5] * pb[(key1 >> 8) * 4];
        f0 = mI < n ? f0 : 0.0;
        f1 = mI + 1 < n ? f1 : 0.0;
        if (TORQUE) {
            a0 = __fma_rn(cp[m0], f0, a0);
            b0 = __fma_rn(cp[m1], f1, b0);
        } else {
            a0 = __fma_rn(cp[m0 * 3], f0, a0);
            a1 = __fma_rn(cp[m0 * 3 + 1], f0, a1);
            a2 = __fma_rn(cp[m0 * 3 + 2], f0, a2);
            b0 = __fma_rn(cp[m1 * 3], f1, b0);
            b1 = __fma_rn(cp[m1 * 3 + 1], f1, b1);
            b2 = __fma_rn(cp[m1 * 3 + 2], f1, b2);
        }
    }
    if (mI < nmax) {
        const int m0 = mI < nlast ? mI : nlast;
        const unsigned key0 = kp[m0];
        double f0 = pa[(key0 & 255u) * 5] * pb[(key0 >> 8) * 4];
        f0 = mI < n ? f0 : 0.0;
        if (TORQUE) {
            a0 = __fma_rn(cp[m0], f0, a0);
        } else {
            a0 = __fma_rn(cp[m0 * 3], f0, a0);
            a1 = __fma_rn(cp[m0 * 3 + 1], f0, a1);
            a2 = __fma_rn(cp[m0 * 3 + 2], f0, a2);
        }
    }
    a0 += b0;
    a1 += b1;
    a2 += b2;
}

__global__ void __launch_bounds__(K3_THREADS, 2)
k_constraints(Batch B, const double* __restrict__ kin, double* __restrict__ g, double* __restrict__ jac, int cand_arena) {
    extern __shared__ __align__(16) unsigned char k3_raw[];
    K3Smem& S = *reinterpret_cast<K3Smem*>(k3_raw);
    const int NJ = B.NJ, O = B.O, T = B.T;
    const int m = B.m();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rows = NJ * TB * O;
    const unsigned meta_bytes = r16(unsigned(rows) * 4u);
    unsigned char* const s_meta0 = k3_raw + sizeof(K3Smem);
    double* const s_cand0 = reinterpret_cast<double*>(k3_raw + sizeof(K3Smem) + 2 * meta_bytes);
    const int cpp = T / TB;  // chunks per problem
    const long long nchunks = (long long)B.nprob * cpp;
    const int c_begin = int(nchunks * blockIdx.x / gridDim.x), c_end = int(nchunks * (blockIdx.x + 1) / gridDim.x);
    const int nmine = c_end - c_begin;
    if (nmine <= 0) return;
    const int ntab = TB * (NF + NJ);
    const float invO = O > 0 ? 1.0f / float(O) : 0.0f;
    const int rec_per_chunk = int(B.hp_chunk_records());
    const int n_slice_items = (NF + NJ) * (TB / 4), n_row_tiles = (rows + 31) >> 5;
    const int I_ISSUE = n_slice_items / 2;                  // position of the "place and issue TAB(j+1)" item
    const int I_TORQUE = n_slice_items + 1 + n_row_tiles;   // torque rows of chunk j-1
    const int ITEMS = n_slice_items + n_row_tiles + 2;

    auto problem_of = [&](int c) { return B.plist ? B.plist[c / cpp] : c / cpp; };
    auto issue_g1 = [&](int c) {  // one thread
        const int s = (c - c_begin) & 1;
        const int p = problem_of(c), tb = c % cpp;
        const size_t t0 = size_t(p) * T + size_t(tb) * TB;
        K3Stage& Q = S.st[s];
        const unsigned b_nl = unsigned(TB * NJ) * 4u, b_nu = unsigned(TB * NF) * 4u, b_l = unsigned(TB * NJ) * 24u,
                       b_u = unsigned(TB * NF) * 8u;
        mbar_arrive_expect_tx(&S.bar_g1[s], b_nl + b_nu + 2 * b_l + 2 * b_u);
        bulk_g2s(Q.nl, B.link_n + t0 * NJ, b_nl, &S.bar_g1[s]);
        bulk_g2s(Q.nu, B.u_n + t0 * NF, b_nu, &S.bar_g1[s]);
        bulk_g2s(Q.cen_l, B.link_c + t0 * NJ * 3, b_l, &S.bar_g1[s]);
        bulk_g2s(Q.rad_l, B.link_r + t0 * NJ * 3, b_l, &S.bar_g1[s]);
        bulk_g2s(Q.cen_u, B.u_c + t0 * NF, b_u, &S.bar_g1[s]);
        bulk_g2s(Q.rad_u, B.u_r + t0 * NF, b_u, &S.bar_g1[s]);
    };
    auto issue_row = [&](int c, int total) {  // one thread: row words + candidate records in use
        const int s = (c - c_begin) & 1;
        const int p = problem_of(c), tb = c % cpp;
        int nrec = total < rec_per_chunk ? total : rec_per_chunk;
        nrec = nrec < cand_arena ? nrec : cand_arena;
        S.rw[s].staged = nrec;
        const unsigned b_meta = unsigned(rows) * 4u;
        mbar_arrive_expect_tx(&S.bar_row[s], b_meta + unsigned(nrec) * 32u);
        if (b_meta) bulk_g2s(s_meta0 + size_t(s) * meta_bytes, B.hp_meta + (size_t(p) * cpp + tb) * rows, b_meta, &S.bar_row[s]);
        if (nrec > 0)
            bulk_g2s(s_cand0 + size_t(s) * cand_arena * 4, B.hp_cand + (size_t(p) * cpp + tb) * size_t(rec_per_chunk) * 4,
                     unsigned(nrec) * 32u, &S.bar_row[s]);
    };
    // TAB(c): one warp, after G1(c) has landed.  Tables task-major (q = task*TB + tt, tasks = 7 joints, then NJ links),
    // two per lane, packed in the arena in that order
    auto issue_tab = [&](int c) {
        const int s = (c - c_begin) & 1;
        const int p = problem_of(c), tb = c % cpp;
        const size_t t0 = size_t(p) * T + size_t(tb) * TB;
        K3Stage& Q = S.st[s];
        unsigned kb[2], cb[2], bytes[2];
        const unsigned short* ksrc[2];
        const double* csrc[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int q = lane + 32 * h;
            kb[h] = cb[h] = 0;
            ksrc[h] = nullptr;
            csrc[h] = nullptr;
            if (q < ntab) {
                const int task = q / TB, tt = q % TB;
                if (task < NF) {
                    const int n = Q.nu[tt * NF + task];
                    const size_t idx = (t0 + tt) * NF + task;
                    kb[h] = r16(unsigned(n) * 2u);
                    cb[h] = r16(unsigned(n) * 8u);
                    ksrc[h] = B.u_key + idx * B.capU;
                    csrc[h] = B.u_g + idx * B.capU;
                } else {
                    const int l = task - NF;
                    const int n = Q.nl[tt * NJ + l];
                    const size_t idx = (t0 + tt) * NJ + l;
                    kb[h] = r16(unsigned(n) * 2u);
                    cb[h] = r16(unsigned(n) * 24u);
                    ksrc[h] = B.link_key + idx * B.capL;
                    csrc[h] = B.link_g + idx * B.capL * 3;
                }
            }
            bytes[h] = kb[h] + cb[h];
        }
        unsigned inc0 = bytes[0], inc1 = bytes[1];  // exclusive prefix over the 64 slots
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned x0 = __shfl_up_sync(0xffffffffu, inc0, d), x1 = __shfl_up_sync(0xffffffffu, inc1, d);
            if (lane >= d) {
                inc0 += x0;
                inc1 += x1;
            }
        }
        const unsigned tot0 = __shfl_sync(0xffffffffu, inc0, 31);
        const unsigned off[2] = {inc0 - bytes[0], tot0 + inc1 - bytes[1]};
        unsigned tx = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int q = lane + 32 * h;
            if (q < ntab) {
                const bool staged = off[h] + bytes[h] <= unsigned(K3_TABLE_ARENA);
                Q.toff[q] = staged ? int(off[h]) : -1;
                if (staged) tx += bytes[h];
                else bytes[h] = 0;
            } else {
                bytes[h] = 0;
            }
        }
        mbar_arrive_expect_tx(&S.bar_tab[s], tx);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (bytes[h]) {  // keys first, then the coefficients (both 16-byte aligned)
                bulk_g2s(Q.arena + off[h], ksrc[h], kb[h], &S.bar_tab[s]);
                bulk_g2s(Q.arena + off[h] + kb[h], csrc[h], cb[h], &S.bar_tab[s]);
            }
        }
    };
    auto total_of = [&](int c) { return O > 0 ? *B.hp_total_of(problem_of(c), c % cpp) : 0; };

    if (tid == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&S.bar_g1[s], 1);
            mbar_init(&S.bar_tab[s], 32);
            mbar_init(&S.bar_row[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.ticket = 0;
        for (int i = 0; i < 4; i++) S.zero_d[i] = 0.0;
        for (int i = 0; i < 8; i++) S.zero_k[i] = 0;
    }
    __syncthreads();
    int next_total = 0;  // (thread 0) candidate records in use of the chunk whose ROW copy is issued next
    if (tid == 0) {
        issue_g1(c_begin);
        next_total = total_of(c_begin);
    }
    if (warp == 0) {
        mbar_wait(&S.bar_g1[0], 0);
        issue_tab(c_begin);
    }
    int cur_p = -1;
    const int grp = lane >> 3, v = lane & 7;
    const double* const pa = &S.pwA[0][v <= 4 ? v : 0];
    const double* const pb = &S.pwB[0][v >= 5 ? v - 4 : 0];
    int ticket = -1;  // warp-uniform: the item this warp holds

    // iteration j: slices of chunk c_begin + j (j < nmine) and rows of chunk c_begin + j - 1 (j >= 1)
    for (int j = 0; j <= nmine; j++) {
        const int c = c_begin + j;
        const bool has_slices = j < nmine, has_rows = j >= 1;
        const int s = j & 1;
        if (tid == 0) {
            if (j + 1 < nmine) issue_g1(c + 1);
            if (has_slices) {
                issue_row(c, next_total);
                if (j + 1 < nmine) next_total = total_of(c + 1);
            }
        }
        const int p = has_slices ? problem_of(c) : cur_p;
        if (has_slices && p != cur_p) {
            // power-product tables of this problem's k.  The rows of chunk j-1 (previous problem) do not read them.
            cur_p = p;
            if (tid < NF) S.k[tid] = kin[size_t(p) * NF + tid];
            __syncthreads();
            if (tid < 256) {  // table A: entry tid, variables k0..k3
                double f[4], d[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double k = S.k[q];
                    const int dg = (tid >> (2 * q)) & 3;
                    f[q] = dg == 0 ? 1.0 : (dg == 1 ? k : (dg == 2 ? k * k : k * k * k));
                    d[q] = dg == 0 ? 0.0 : (dg == 1 ? 1.0 : (dg == 2 ? 2.0 * k : 3.0 * (k * k)));
                }
                const double p01 = f[0] * f[1], p23 = f[2] * f[3];
                double* row = S.pwA[tid];
                row[0] = p01 * p23;
                row[1] = d[0] * f[1] * p23;
                row[2] = f[0] * d[1] * p23;
                row[3] = p01 * (d[2] * f[3]);
                row[4] = p01 * (f[2] * d[3]);
            }
            if (tid < 64) {  // table B: entry tid, variables k4..k6
                double f[3], d[3];
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const double k = S.k[4 + q];
                    const int dg = (tid >> (2 * q)) & 3;
                    f[q] = dg == 0 ? 1.0 : (dg == 1 ? k : (dg == 2 ? k * k : k * k * k));
                    d[q] = dg == 0 ? 0.0 : (dg == 1 ? 1.0 : (dg == 2 ? 2.0 * k : 3.0 * (k * k)));
                }
                double* row = S.pwB[tid];
                row[0] = f[0] * f[1] * f[2];
                row[1] = d[0] * f[1] * f[2];
                row[2] = f[0] * d[1] * f[2];
                row[3] = f[0] * f[1] * d[2];
            }
            if (tid == 0) {
                bool in = true;
                for (int q = 0; q < NF; q++) in = in && (fabs(S.k[q]) <= K_DOMAIN);
                S.rw[s].in_domain = in ? 1 : 0;
            }
            __syncthreads();
        } else if (has_slices && tid == 0) {
            S.rw[s].in_domain = S.rw[s ^ 1].in_domain;  // same problem as the previous chunk
        }

        // slice side: chunk c, stage s.  row side: chunk c - 1, stage s ^ 1
        K3Stage& Q = S.st[s];
        K3Rows& RS = S.rw[s];
        const K3Rows& RR = S.rw[s ^ 1];
        const int tb = c % cpp;
        const size_t t0 = size_t(p) * T + size_t(tb) * TB;
        const bool failed = has_slices && B.status[p] != 0;
        const int pr = has_rows ? problem_of(c - 1) : 0, tbr = (c - 1) % cpp;
        const bool failed_r = has_rows && B.status[pr] != 0;
        double* gpr = g ? g + size_t(pr) * m : nullptr;
        double* jpr = jac ? jac + size_t(pr) * m * NF : nullptr;
        const unsigned* s_meta = reinterpret_cast<const unsigned*>(s_meta0 + size_t(s ^ 1) * meta_bytes);
        const double* s_cand = s_cand0 + size_t(s ^ 1) * cand_arena * 4;
        bool tab_ready = false, row_ready = false;

        for (;;) {
            if (ticket < 0) {
                int t = 0;
                if (lane == 0) t = atomicAdd(&S.ticket, 1);
                ticket = __shfl_sync(0xffffffffu, t, 0);
            }
            if (ticket >= (j + 1) * ITEMS) break;  // an item of the next iteration: keep it, go to the barrier
            const int item = ticket - j * ITEMS;
            ticket = -1;

            if (item < n_slice_items + 1 && item != I_ISSUE) {
                // ---- slice item: the tables of four consecutive intervals of one joint / link of chunk c
                if (!has_slices) continue;
                const int task = item < I_ISSUE ? item : item - 1;
                if (!tab_ready) {
                    mbar_wait(&S.bar_tab[s], unsigned(j >> 1) & 1u);
                    tab_ready = true;
                }
                if (failed) continue;
                const int sub = task % (TB / 4), which = task / (TB / 4);  // longest tables (torques) first
                const int tt = sub * 4 + grp;
                const bool torque = which < NF;
                const int l = which - NF;
                const int n = torque ? Q.nu[tt * NF + which] : Q.nl[tt * NJ + l];
                const int to = Q.toff[which * TB + tt];
                int nmax = n;
                nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 8));
                nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 16));
                const bool all_staged = __all_sync(0xffffffffu, to >= 0);
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                if (all_staged) {
                    const unsigned short* kp = n > 0 ? reinterpret_cast<const unsigned short*>(Q.arena + to) : S.zero_k;
                    const double* cp = n > 0 ? reinterpret_cast<const double*>(Q.arena + to + r16(unsigned(n) * 2u)) : S.zero_d;
                    if (torque) slice_group<true>(kp, cp, n, nmax, pa, pb, a0, a1, a2);
                    else slice_group<false>(kp, cp, n, nmax, pa, pb, a0, a1, a2);
                } else {  // a table that did not fit the arena: same walk over global memory
                    const size_t idx = torque ? (t0 + tt) * NF + which : (t0 + tt) * NJ + l;
                    const unsigned short* kp = torque ? B.u_key + idx * B.capU : B.link_key + idx * B.capL;
                    const double* cp = torque ? B.u_g + idx * B.capU : B.link_g + idx * B.capL * 3;
                    if (n == 0) {
                        kp = S.zero_k;
                        cp = S.zero_d;
                    }
                    if (torque) slice_group<true>(kp, cp, n, nmax, pa, pb, a0, a1, a2);
                    else slice_group<false>(kp, cp, n, nmax, pa, pb, a0, a1, a2);
                }
                if (torque) {
                    const int i = tt * NF + which;
                    if (v == 0) {
                        const double value = Q.cen_u[i] + a0;
                        const double r = Q.rad_u[i];
                        RS.tg[i] = ((value - r) + (value + r)) * 0.5;  // centre of Interval(c - r, c + r) (KPR/NLPclass.cu:306)
                    } else {
                        RS.tj[i * NF + (v - 1)] = a0;
                    }
                } else {
                    const int i = tt * NJ + l;
                    if (v == 0) {
                        const double acc[3] = {a0, a1, a2};
#pragma unroll
                        for (int e = 0; e < 3; e++) {
                            const double value = Q.cen_l[i * 3 + e] + acc[e];
                            const double r = Q.rad_l[i * 3 + e];
                            RS.lc[tt][l][e] = ((value - r) + (value + r)) * 0.5;  // getCenter(slice()) (KPR/NLPclass.cu:313)
                        }
                    } else {
                        RS.dlc[tt][l][v - 1][0] = a0;
                        RS.dlc[tt][l][v - 1][1] = a1;
                        RS.dlc[tt][l][v - 1][2] = a2;
                    }
                }
            } else if (item == I_ISSUE) {
                // ---- place and fetch the tables of chunk c + 1 (its G1 was issued at the top of this iteration)
                if (j + 1 < nmine) {
                    mbar_wait(&S.bar_g1[s ^ 1], unsigned((j + 1) >> 1) & 1u);
                    issue_tab(c + 1);
                }
            } else if (item == I_TORQUE) {
                // ---- torque rows of chunk c - 1: contiguous runs of g and of the Jacobian; Bezier rows once per problem
                if (!has_rows) continue;
                if (failed_r) {
                    // the build of this problem overflowed a table (ARMOUR_ERR_CAPACITY): its reach sets are not valid.
                    // Fail-safe rows: every torque and collision row violated, zero Jacobian.
                    for (int i = lane; i < TB * NF; i += 32) {
                        if (gpr) gpr[size_t(tbr) * TB * NF + i] = 1e300;
                        if (jpr)
                            for (int q = 0; q < NF; q++) jpr[(size_t(tbr) * TB * NF + i) * NF + q] = 0.0;
                    }
                    for (int x = lane; x < rows; x += 32) {
                        const int o = x % O, ltt = x / O, tt = ltt % TB, l = ltt / TB;
                        const size_t r = size_t(NF) * T + (size_t(l) * T + tbr * TB + tt) * O + o;
                        if (gpr) gpr[r] = 1e300;
                        if (jpr)
                            for (int q = 0; q < NF; q++) jpr[r * NF + q] = 0.0;
                    }
                } else {
                    if (gpr)
                        for (int i = lane; i < TB * NF; i += 32) gpr[size_t(tbr) * TB * NF + i] = RR.tg[i];
                    if (jpr)
                        for (int i = lane; i < TB * NF * NF; i += 32) jpr[size_t(tbr) * TB * NF * NF + i] = RR.tj[i];
                    if (pr == 0 && B.link_sliced)
                        for (int i = lane; i < TB * NJ * 3; i += 32)
                            B.link_sliced[(size_t(tbr) * TB * NJ) * 3 + i] = RR.lc[i / (NJ * 3)][(i / 3) % NJ][i % 3];
                }
                if (tbr == 0 && lane < NF) {  // Bezier joint-limit rows (KPR/Trajectory.cu:256-540)
                    const int i = lane;
                    const double D = c_robot.duration;
                    const double q0 = B.q0[size_t(pr) * NF + i];
                    const double a = B.qd0[size_t(pr) * NF + i] * D;
                    const double b = B.qdd0[size_t(pr) * NF + i] * D * D;
                    const double kn = kin[size_t(pr) * NF + i];
                    const int off = NF * T + NJ * T * O;
                    for (int vel = 0; vel < 2; vel++) {
                        double mn, mx, dmn, dmx;
                        bez_extrema(vel == 1, q0, a, b, c_robot.k_range[i], D, kn, &mn, &mx, &dmn, &dmx);
                        const int r0 = off + vel * 2 * NF + i;
                        if (gpr) {
                            gpr[r0] = failed_r ? 0.0 : mn;
                            gpr[r0 + NF] = failed_r ? 0.0 : mx;
                        }
                        if (jpr) {
                            for (int q = 0; q < NF; q++) {
                                jpr[size_t(r0) * NF + q] = (q == i && !failed_r) ? dmn : 0.0;
                                jpr[size_t(r0 + NF) * NF + q] = (q == i && !failed_r) ? dmx : 0.0;
                            }
                        }
                    }
                }
            } else {
                // ---- row tile of chunk c - 1: rows x = (l*TB + tt)*O + o, one lane each
                if (!has_rows || failed_r || O == 0 || !RR.in_domain) continue;
                if (!row_ready) {
                    mbar_wait(&S.bar_row[s ^ 1], unsigned((j - 1) >> 1) & 1u);
                    row_ready = true;
                }
                const int x = (item - (n_slice_items + 1)) * 32 + lane;
                if (x >= rows) continue;
                const int ltt = __float2int_rz((float(x) + 0.5f) * invO);  // x / O (exact: x < 2^22)
                const int o = x - ltt * O;
                const int tt = ltt % TB, l = ltt / TB;
                const unsigned meta = s_meta[x];
                const int n = int(meta & 255u), off = int(meta >> 8);
                if (n == HP_OVERFLOW) continue;  // no stored list: k_constraints_slow writes this row
                const int row_i = NF * T + (l * T + tbr * TB + tt) * O + o;
                const double c0 = RR.lc[tt][l][0], c1 = RR.lc[tt][l][1], c2 = RR.lc[tt][l][2];
                const double2* rec = reinterpret_cast<const double2*>(
                    off + n <= RR.staged ? s_cand + size_t(off) * 4
                                         : B.hp_cand + ((size_t(pr) * cpp + tbr) * size_t(rec_per_chunk) + off) * 4);
                double max_elt = -100000000;
                double A0 = 0, A1 = 0, A2 = 0;  // minus the winning signed normal
                for (int q = 0; q < n; q++) {
                    const double2 u = rec[2 * q], w = rec[2 * q + 1];
                    const double val = (u.x * c0 + u.y * c1 + w.x * c2) - w.y;
                    if (val > max_elt) {  // strict '>' in scan order: KPR/CollisionChecking.cu:264-276
                        max_elt = val;
                        A0 = -u.x; A1 = -u.y; A2 = -w.x;
                    }
                }
                if (gpr) gpr[row_i] = -max_elt;
                if (jpr) {
                    const double2* dk = reinterpret_cast<const double2*>(&RR.dlc[tt][l][0][0]);
                    double* out = jpr + size_t(row_i) * NF;
#pragma unroll
                    for (int q = 0; q < NF; q++) {
                        const double2 xy = dk[2 * q];
                        const double z = dk[2 * q + 1].x;
                        // -(C.dk) for a 'pos' winner, +(C.dk) for 'neg' (:286-295); the sign is folded into A.  Fused like
                        // the reference's own kernel (nvcc contracts max_A_elt.dot(dk)); 56 contiguous bytes per lane
                        out[q] = __fma_rn(A0, xy.x, __fma_rn(A1, xy.y, A2 * z));
                    }
                }
            }
        }
        if (j == nmine && has_rows && !row_ready) mbar_wait(&S.bar_row[s ^ 1], unsigned((j - 1) >> 1) & 1u);  // nothing in flight at exit
        __syncthreads();
    }
}

// k_constraints_slow: the collision rows k_constraints leaves out — rows without a stored candidate list (more than HP_CAP
// survivors, or no room left in the chunk) and every row of a problem whose k lies outside the box the lists were built
// for.  One CTA per problem, which returns at once in the usual case (no such row); otherwise one thread per row slices
// the link reach set in the reference's factor order (KPR/PZsparse.cu:404-555) and scans all 72 half-spaces computed from
// the generators (KPR/CollisionChecking.cu:169-299).
__global__ void __launch_bounds__(128)
k_constraints_slow(Batch B, const double* __restrict__ kin, double* __restrict__ g, double* __restrict__ jac) {
    const int p = B.plist ? B.plist[blockIdx.x] : int(blockIdx.x);
    const int NJ = B.NJ, O = B.O, T = B.T;
    if (O == 0 || B.status[p] != 0) return;
    __shared__ double2 kpd[NF][4];  // {k_j^d, d/dk_j k_j^d} for d = 0..3
    bool in = true;
    for (int j = 0; j < NF; j++) in = in && (fabs(kin[size_t(p) * NF + j]) <= K_DOMAIN);
    if (in && *B.hp_slow_of(p) == 0) return;
    if (threadIdx.x < NF) {
        const double k = kin[size_t(p) * NF + threadIdx.x];
        kpd[threadIdx.x][0] = make_double2(1.0, 0.0);
        kpd[threadIdx.x][1] = make_double2(k, 1.0);
        kpd[threadIdx.x][2] = make_double2(k * k, 2.0 * k);
        kpd[threadIdx.x][3] = make_double2(k * k * k, 3.0 * (k * k));
    }
    __syncthreads();
    const int m = B.m();
    const int per_chunk = NJ * TB * O;
    double* gp = g ? g + size_t(p) * m : nullptr;
    double* jp = jac ? jac + size_t(p) * m * NF : nullptr;
    for (int r = threadIdx.x; r < per_chunk * (T / TB); r += blockDim.x) {
        const int tb = r / per_chunk, x = r - tb * per_chunk;
        const size_t chunk = size_t(p) * (T / TB) + tb;
        if (in && (B.hp_meta[chunk * per_chunk + x] & 255u) != HP_OVERFLOW) continue;
        const int o = x % O, ltt = x / O, tt = ltt % TB, l = ltt / TB;
        const size_t idx = (size_t(p) * T + tb * TB + tt) * NJ + l;
        double c[3], dk[NF][3];
        const int n = B.link_n[idx];
        for (int e = 0; e < 3; e++) {
            double value = B.link_c[idx * 3 + e];
            double grad[NF] = {0, 0, 0, 0, 0, 0, 0};
            for (int mI = 0; mI < n; mI++) {
                const unsigned key = B.link_key[idx * B.capL + mI];
                double val = B.link_g[(idx * B.capL + mI) * 3 + e];
                double D[NF];
#pragma unroll
                for (int j = 0; j < NF; j++) {
                    const double2 fd = kpd[j][(key >> (2 * j)) & 3];
                    D[j] = val * fd.y;
#pragma unroll
                    for (int v = 0; v < j; v++) D[v] *= fd.x;
                    val *= fd.x;
                }
                value += val;
#pragma unroll
                for (int v = 0; v < NF; v++) grad[v] += D[v];
            }
            const double rr = B.link_r[idx * 3 + e];
            c[e] = ((value - rr) + (value + rr)) * 0.5;
#pragma unroll
            for (int v = 0; v < NF; v++) dk[v][e] = grad[v];
        }
        double r_max, A0, A1, A2;
        row_from_generators(B.obstacles + (size_t(p) * O + o) * 12, B.link_gens + idx * 18, c[0], c[1], c[2], &r_max, &A0, &A1, &A2);
        const size_t row = size_t(NF) * T + (size_t(l) * T + tb * TB + tt) * O + o;
        if (gp) gp[row] = -r_max;
        if (jp)
            for (int v = 0; v < NF; v++) jp[row * NF + v] = A0 * dk[v][0] + A1 * dk[v][1] + A2 * dk[v][2];
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
    cudaError_t e = cudaMemsetAsync(B.hp_total, 0, size_t(B.nprob) * (B.T / TB + 1) * sizeof(int), st);
    if (e != cudaSuccess) return e;
    const int rows = B.NJ * B.T * B.O;
    dim3 grid((rows + HP_ROWS - 1) / HP_ROWS, B.nprob);
    k_hyperplanes<<<grid, HP_THREADS, 0, st>>>(B);
    return cudaGetLastError();
}

// Shared memory of one persistent k_constraints CTA: the fixed part plus, per chunk parity, the row words and a
// candidate arena sized for the chunk's expected 2.5 records per row — bounded so that two CTAs stay resident per SM
// (else one).
constexpr int K3_SMEM_MAX = 220 * 1024;
inline void k3_smem_plan(const Batch& B, int* smem_bytes, int* cand_arena, int* ctas_per_sm) {
    const int rows = B.chunk_rows();
    const int fixed = int(sizeof(K3Smem)) + 2 * int((size_t(rows) * 4 + 15) & ~size_t(15));
    int want = rows * 5 / 2 + 32;
    if (want > int(B.hp_chunk_records())) want = int(B.hp_chunk_records());
    const int budgets[2] = {112 * 1024, K3_SMEM_MAX};
    int rec = 0, b = 0;
    for (b = 0; b < 2; b++) {
        rec = (budgets[b] - fixed) / 64;  // two stages of 32-byte records
        if (rec >= want || b == 1) break;
    }
    if (rec < 0) rec = 0;
    if (rec > want) rec = want;
    *cand_arena = rec;
    *smem_bytes = fixed + 2 * rec * 32;
    *ctas_per_sm = (b == 0) ? 2 : 1;
}
cudaError_t launch_constraints(const Batch& B, const double* d_k, double* d_g, double* d_jac, cudaStream_t st) {
    if (B.nprob == 0) return cudaSuccess;
    int smem = 0, arena = 0, per_sm = 1;
    k3_smem_plan(B, &smem, &arena, &per_sm);
    if (smem > K3_SMEM_MAX) return cudaErrorInvalidValue;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const long long nchunks = (long long)B.nprob * (B.T / TB);
    const int grid = int(nchunks < (long long)sms * per_sm ? nchunks : (long long)sms * per_sm);
    k_constraints<<<grid, K3_THREADS, smem, st>>>(B, d_k, d_g, d_jac, arena);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || B.O == 0) return e;
    k_constraints_slow<<<B.nprob, 128, 0, st>>>(B, d_k, d_g, d_jac);
    return cudaGetLastError();
}
cudaError_t launch_verdict(const Batch& B, const double* d_g, int* d_feasible, int* d_first, cudaStream_t st) {
    if (B.nprob == 0) return cudaSuccess;
    k_verdict<<<B.nprob, 256, 0, st>>>(B, d_g, d_feasible, d_first);
    return cudaGetLastError();
}

}  // namespace armour
