// K1/K2: reach-set construction on the device.  One CTA builds ALL reach sets of one
// (planning problem, time interval): the joint reachable set of the Bezier trajectory (K2, prologue),
// the PZ forward kinematics of the link volumes, one pass of PZ recursive Newton-Euler carrying the
// nominal and the interval-parameter radius lanes, and the robust-input radius (epilogue).
//
// Replaces, per time interval, BezierCurve::makePolyZono (KPR/Trajectory.cu:63-254),
// KinematicsDynamics::fk / rnea_nominal / rnea_interval (KPR/Dynamics.cu:69-181), reduce_link_PZ / reduce
// (KPR/PZsparse.cu:352-402) and sections II.B-II.C of main() (KPR/armour_main.cu:96-210), which the
// reference runs on the host under OpenMP.  Operation order and every simplify() point follow the
// reference (see k1_pz.cuh for how a simplify() is carried out here).
//
// CTAs are persistent: each takes (problem, interval) units from a global counter, so a batch of
// worlds x replans fills the 148 SMs without any cross-CTA communication.
#pragma once
#ifndef ARMOUR_EMU
#include "../../include/armour_b200.h"
#endif
#include "bezier.cuh"
#include "device_constants.cuh"
#include "k1_interval.cuh"
#include "k1_pz.cuh"
#include "layout.h"

namespace armour {
namespace k1 {

// fixed shared-memory region for the joint reachable set of the current interval
constexpr int ROT_WORDS = 27 + 3 + 27;  // 3x3 PZ with at most 3 monomials
constexpr int SCL_WORDS = 3 + 2 + 2;    // scalar PZ with at most 2 monomials
constexpr int JRS_WORDS = (MAXJ + 1) * ROT_WORDS + 3 * NF * SCL_WORDS;
constexpr double RADIUS_SLACK = 1.0 + 0x1p-40;  // outward slack on exported radii (see DESIGN.md "soundness")

struct K1Params {
    Batch B;
    int* work;          // global unit counter (reset before every launch)
    double* gscr;       // [grid][gscr_words]
    int gscr_words;     // = fn_words (F / N blocks of one unit) + spill space of the arena
    int fn_words;
    char* gtab;         // [grid][gtab_bytes], all zero between launches
    int gtab_bytes;
    int arena_words;    // shared-memory arena per CTA
    int tab_s_bytes;    // shared-memory table pool per CTA
    const int* units;   // optional explicit unit list (p*T + t); nullptr = all units of the batch
    int nunits;
    int* stats;         // [4]: max arena words used, tables placed in global memory, failed units, units done
};

struct Jrs {
    PZH R[MAXJ + 1];
    PZH qd[NF], qda[NF], qdda[NF];
};

K1_DI void indep_range(double v_lb, double v_ub, double s_lb, double s_ub, double e1s, double e1v, double e2s,
                       double e2v, double* radius, double* center) {  // KPR/Trajectory.cu:80-94
    double lb = v_lb, ub = v_ub;
    if (lb > ub) {
        const double x = lb;
        lb = ub;
        ub = x;
    }
    if (s_lb < e1s && e1s < s_ub) {
        lb = fmin(lb, e1v);
        ub = fmax(ub, e1v);
    }
    if (s_lb < e2s && e2s < s_ub) {
        lb = fmin(lb, e2v);
        ub = fmax(ub, e2v);
    }
    *radius = (ub - lb) * 0.5;
    *center = (lb + ub) * 0.5;
}

// scalar PZ  c + coef0 * x_{key0} + coef1 * x_{key1}, simplified (KPR/PZsparse.cu:120-136)
K1_DI void write_scalar_pz(double* p, int* n_out, double thr, double center, double c0, u64 k0, double c1, u64 k1_) {
    double rad = 0.0;
    int n = 0;
    u64 keys[2];
    double cf[2];
    const double cc[2] = {c0, c1};
    const u64 kk[2] = {k0, k1_};
    for (int i = 0; i < 2; i++) {
        if (sqrt(cc[i] * cc[i]) <= thr) {
            rad = __dadd_ru(rad, fabs(cc[i]));
        } else {
            keys[n] = kk[i];
            cf[n] = cc[i];
            n++;
        }
    }
    p[0] = center;
    p[1] = rad;
    p[2] = rad;
    u64* pk = reinterpret_cast<u64*>(p + 3);
    for (int i = 0; i < n; i++) pk[i] = keys[i];
    for (int i = 0; i < n; i++) p[3 + n + i] = cf[i];
    *n_out = n;
}

// Joint reachable set of joint i over interval t (KPR/Trajectory.cu:15-61 for the extrema of the
// k-independent parts, :63-254 for the interval itself).  Executed by one thread per joint.
K1_OP void jrs_joint(int i, int t, int T, double q0, double qd0, double qdd0, double thr, double* rot_blk, int* rot_n,
                     double* qd_blk, int* qd_n, double* qda_blk, int* qda_n, double* qdda_blk, int* qdda_n) {
    const RobotConstants& rc = c_robot;
    const double D = rc.duration;
    const double a = qd0 * D, b = qdd0 * D * D;
    const double kr = rc.k_range[i];
    const double ds = 1.0 / T;
    const double s_lb = t * ds, s_ub = (t + 1) * ds;
    // interior extrema of the k-independent parts (BezierCurve ctor)
    double es[3][2], ev[3][2];
    {
        const double sq = sqrt(64 * pw2(a) + 14 * a * b + pw2(b));
        es[0][0] = (2 * a + b + sq) / (5 * (6 * a + b));
        es[0][1] = (2 * a + b - sq) / (5 * (6 * a + b));
        ev[0][0] = bez_q_indep(q0, a, b, es[0][0]);
        ev[0][1] = bez_q_indep(q0, a, b, es[0][1]);
    }
    {
        const double sq = sqrt(6 * (54 * pw2(a) + 14 * a * b + pw2(b)));
        es[1][0] = (18 * a + 4 * b + sq) / (10 * (6 * a + b));
        es[1][1] = (18 * a + 4 * b - sq) / (10 * (6 * a + b));
        ev[1][0] = bez_qd_indep(a, b, es[1][0], D);
        ev[1][1] = bez_qd_indep(a, b, es[1][1], D);
    }
    {
        const double sq = sqrt(2 * (152 * pw2(a) + 42 * a * b + 3 * pw2(b)));
        es[2][0] = (32 * a + 6 * b + sq) / (10 * (6 * a + b));
        es[2][1] = (32 * a + 6 * b - sq) / (10 * (6 * a + b));
        ev[2][0] = bez_qdd_indep(a, b, es[2][0], D);
        ev[2][1] = bez_qdd_indep(a, b, es[2][1], D);
    }

    // Part 1: position -> cos / sin Taylor models (:75-134)
    double kd_lb = pw3(s_lb) * (6 * pw2(s_lb) - 15 * s_lb + 10);
    double kd_ub = pw3(s_ub) * (6 * pw2(s_ub) - 15 * s_ub + 10);
    double kd_center = (kd_ub + kd_lb) * 0.5;
    double kd_radius = (kd_ub - kd_lb) * 0.5 * kr;
    double ki_radius, q_c;
    indep_range(bez_q_indep(q0, a, b, s_lb), bez_q_indep(q0, a, b, s_ub), s_lb, s_ub, es[0][0], ev[0][0], es[0][1],
                ev[0][1], &ki_radius, &q_c);
    const Itv qr = iv(-kd_radius - ki_radius - rc.qe, kd_radius + ki_radius + rc.qe);
    const Itv kint = iv(-kr, kr);
    const double sq_c = sin(q_c), cq_c = cos(q_c);
    const Itv arg = iv_add(iv_addd(q_c, iv_muld(kd_center, kint)), qr);
    const Itv e2 = iv_pow2(iv_add(qr, iv_muld(kd_center, kint)));
    Itv cos_r = iv_sub(iv_muld(sq_c, iv_neg(qr)), iv_mul(iv_muld(0.5, iv_cos(arg)), e2));
    double cos_c = cq_c + iv_mid(cos_r);
    cos_r = iv_subd(cos_r, iv_mid(cos_r));
    const double cos_k = -kd_center * kr * sq_c, cos_e = iv_rad(cos_r);
    Itv sin_r = iv_sub(iv_muld(cq_c, qr), iv_mul(iv_muld(0.5, iv_sin(arg)), e2));
    double sin_c = sq_c + iv_mid(sin_r);
    sin_r = iv_subd(sin_r, iv_mid(sin_r));
    const double sin_k = kd_center * kr * cq_c, sin_e = iv_rad(sin_r);

    // 3x3 rotation about z from the cos / sin models (KPR/PZsparse.cu:179-250), simplified, then
    // R = Rrpy * Rz (KPR/Trajectory.cu:136-144): a product with a constant left operand.
    {
        double rotc[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        rotc[0] = cos_c;
        rotc[3] = -1.0 * sin_c;
        rotc[1] = sin_c;
        rotc[4] = cos_c;
        double rot_rad[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        double M[3][9];
        u64 mk[3];
        int nm = 0;
        for (int which = 0; which < 3; which++) {
            double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (which == 0) {  // the two k_i monomials merged
                m[0] = cos_k;
                m[1] = sin_k;
                m[3] = -0.0 + (-1.0 * sin_k);
                m[4] = cos_k;
            } else if (which == 1) {
                m[0] = cos_e;
                m[3] = -0.0;
                m[4] = cos_e;
            } else {
                m[1] = sin_e;
                m[3] = -1.0 * sin_e;
            }
            if (frobN<9>(m) <= thr) {
                for (int e = 0; e < 9; e++) rot_rad[e] = __dadd_ru(rot_rad[e], fabs(m[e]));
            } else {
                for (int e = 0; e < 9; e++) M[nm][e] = m[e];
                mk[nm] = which == 0 ? key_k(i) : (which == 1 ? key_cosqe(i) : key_sinqe(i));
                nm++;
            }
        }
        const double* Rr = &rc.rrpy[i * 9];
        double cen[9], absR[9], rad[9];
        matmul3<3, false>(Rr, rotc, cen);
        for (int e = 0; e < 9; e++) absR[e] = fabs(Rr[e]);
        matmul3_up<3, false>(absR, rot_rad, rad);
        double outM[3][9];
        u64 outk[3];
        int no = 0;
        for (int j = 0; j < nm; j++) {
            double v[9];
            matmul3<3, false>(Rr, M[j], v);
            if (frobN<9>(v) <= thr) {
                for (int e = 0; e < 9; e++) rad[e] = __dadd_ru(rad[e], fabs(v[e]));
            } else {
                for (int e = 0; e < 9; e++) outM[no][e] = v[e];
                outk[no] = mk[j];
                no++;
            }
        }
        for (int e = 0; e < 9; e++) {
            rot_blk[e] = cen[e];
            rot_blk[9 + e] = rad[e];
            rot_blk[18 + e] = rad[e];
        }
        u64* pk = reinterpret_cast<u64*>(rot_blk + 27);
        for (int j = 0; j < no; j++) pk[j] = outk[j];
        for (int j = 0; j < no; j++)
            for (int e = 0; e < 9; e++) rot_blk[27 + no + j * 9 + e] = outM[j][e];
        *rot_n = no;
    }

    // Part 2: velocity (:151-192)
    kd_lb = (30 * pw2(s_lb) * pw2(s_lb - 1)) / D;
    kd_ub = (30 * pw2(s_ub) * pw2(s_ub - 1)) / D;
    if (kd_ub < kd_lb) {
        const double x = kd_lb;
        kd_lb = kd_ub;
        kd_ub = x;
    }
    kd_center = (kd_ub + kd_lb) * 0.5 * kr;
    kd_radius = (kd_ub - kd_lb) * 0.5 * kr;
    double qd_c;
    indep_range(bez_qd_indep(a, b, s_lb, D), bez_qd_indep(a, b, s_ub, D), s_lb, s_ub, es[1][0], ev[1][0], es[1][1],
                ev[1][1], &ki_radius, &qd_c);
    write_scalar_pz(qd_blk, qd_n, thr, qd_c, kd_center, key_k(i), kd_radius + ki_radius + rc.qde, key_qde(i));
    write_scalar_pz(qda_blk, qda_n, thr, qd_c, kd_center, key_k(i), kd_radius + ki_radius + rc.qdae, key_qdae(i));

    // Part 3: acceleration (:195-244)
    const double kQddMax = 0.5 - sqrt(3.0) / 6, kQddMin = 0.5 + sqrt(3.0) / 6;  // KPR/Trajectory.h:7-8
    const double temp_lb = (60 * s_lb * (2 * pw2(s_lb) - 3 * s_lb + 1)) / D / D;
    const double temp_ub = (60 * s_ub * (2 * pw2(s_ub) - 3 * s_ub + 1)) / D / D;
    if (s_ub <= kQddMax) {
        kd_lb = temp_lb;
        kd_ub = temp_ub;
    } else if (s_lb <= kQddMax) {
        kd_lb = fmin(temp_lb, temp_ub);
        kd_ub = (60 * kQddMax * (2 * pw2(kQddMax) - 3 * kQddMax + 1)) / D / D;
    } else if (s_ub <= kQddMin) {
        kd_lb = temp_ub;
        kd_ub = temp_lb;
    } else if (s_lb <= kQddMin) {
        kd_lb = (60 * kQddMin * (2 * pw2(kQddMin) - 3 * kQddMin + 1)) / D / D;
        kd_ub = fmax(temp_lb, temp_ub);
    } else {
        kd_lb = temp_lb;
        kd_ub = temp_ub;
    }
    kd_center = (kd_ub + kd_lb) * 0.5 * kr;
    kd_radius = (kd_ub - kd_lb) * 0.5 * kr;
    double qdd_c;
    indep_range(bez_qdd_indep(a, b, s_lb, D), bez_qdd_indep(a, b, s_ub, D), s_lb, s_ub, es[2][0], ev[2][0], es[2][1],
                ev[2][1], &ki_radius, &qdd_c);
    write_scalar_pz(qdda_blk, qdda_n, thr, qdd_c, kd_center, key_k(i), kd_radius + ki_radius + rc.qddae, key_qddae(i));
}

// ---- exports --------------------------------------------------------------------------------------
// reduce_link_PZ (KPR/PZsparse.cu:370-402) + the k-only table of one link reach set, sorted by key.
K1_OP void export_link(Ctx& c, const PZH& L, const Batch& B, int p, int t, int l) {
    if (c.fail) return;
    const u64* keys = pz_keys(L);
    const double* cf = pz_coef(L);
    const size_t idx = (size_t(p) * B.T + t) * B.NJ + l;
    double rad[3] = {0, 0, 0};
    int nk = 0;
    for (int m = 0; m < L.n; m++) nk += (keys[m] < KEY_K_ONLY);  // uniform (broadcast reads)
    if (nk > B.capL) {
        set_fail(c, FAIL_LINK_CAP);
        return;
    }
    double* gens = B.link_gens + idx * 18;
    if (c.tid < 18) gens[c.tid] = 0.0;
    __syncthreads();
    for (int m = c.tid; m < L.n; m += NT) {
        const u64 key = keys[m];
        const bool konly = key < KEY_K_ONLY;
        const bool gen = !konly && key < KEY_K_LINKS && (key & KEY_K_MASK) == 0;
        if (konly || gen) {
            int rank = 0;
            for (int q = 0; q < L.n; q++) {
                const u64 kq = keys[q];
                const bool same = konly ? (kq < KEY_K_ONLY) : (kq >= KEY_K_ONLY && kq < KEY_K_LINKS && (kq & KEY_K_MASK) == 0);
                rank += (same && kq < key);
            }
            if (konly) {
                B.link_key[idx * B.capL + rank] = uint16_t(key);
                for (int e = 0; e < 3; e++) B.link_g[(idx * B.capL + rank) * 3 + e] = cf[size_t(m) * 3 + e];
            } else if (rank < 3) {
                for (int e = 0; e < 3; e++) gens[e + rank * 3] = cf[size_t(m) * 3 + e];
            } else {
                for (int e = 0; e < 3; e++) rad[e] = __dadd_ru(rad[e], fabs(cf[size_t(m) * 3 + e]));
            }
        } else {
            for (int e = 0; e < 3; e++) rad[e] = __dadd_ru(rad[e], fabs(cf[size_t(m) * 3 + e]));
        }
    }
    for (int e = 0; e < 3; e++) rad[e] = warp_sum_up(rad[e]);
    if (c.lane == 0)
        for (int e = 0; e < 3; e++) c.red[c.warp * RED_STRIDE + e] = rad[e];
    __syncthreads();
    if (c.tid < 3) {
        const int e = c.tid;
        double v = pz_r(L, 0)[e];
        for (int w = 0; w < NW; w++) v = __dadd_ru(v, c.red[w * RED_STRIDE + e]);
        gens[e + (3 + e) * 3] = __dmul_ru(v, RADIUS_SLACK);
        B.link_c[idx * 3 + e] = pz_c(L)[e];
    }
    if (c.tid == 0) B.link_n[idx] = nk;
    __syncthreads();
}

// u_nom.reduce() (KPR/PZsparse.cu:352-368) + the k-only table of one torque reach set; returns through
// shared memory the radius of the reduced nominal PZ and the disturbance radius
// toInterval(u_nom_int - u_nom) = r_int + r_nom before the reduce (KPR/armour_main.cu:134-141).
K1_OP void export_torque(Ctx& c, const PZH& U, const Batch& B, int p, int t, int j, double* s_unom_r, double* s_dist) {
    if (c.fail) return;
    const u64* keys = pz_keys(U);
    const double* cf = pz_coef(U);
    const size_t idx = (size_t(p) * B.T + t) * NF + j;
    int nk = 0;
    for (int m = 0; m < U.n; m++) nk += (keys[m] < KEY_K_ONLY);
    if (nk > B.capU) {
        set_fail(c, FAIL_TORQUE_CAP);
        return;
    }
    double rad = 0.0;
    for (int m = c.tid; m < U.n; m += NT) {
        const u64 key = keys[m];
        if (key < KEY_K_ONLY) {
            int rank = 0;
            for (int q = 0; q < U.n; q++) rank += (keys[q] < key);  // every smaller key is k-only too
            B.u_key[idx * B.capU + rank] = uint16_t(key);
            B.u_g[idx * B.capU + rank] = cf[m];
        } else {
            rad = __dadd_ru(rad, fabs(cf[m]));
        }
    }
    rad = warp_sum_up(rad);
    if (c.lane == 0) c.red[c.warp * RED_STRIDE] = rad;
    __syncthreads();
    if (c.tid == 0) {
        double v = 0.0;
        for (int w = 0; w < NW; w++) v = __dadd_ru(v, c.red[w * RED_STRIDE]);
        const double r_nom = pz_r(U, 0)[0], r_int = pz_r(U, 1)[0];
        const double reduced = __dadd_ru(r_nom, v);
        s_unom_r[j] = reduced;
        s_dist[j] = __dadd_ru(r_int, r_nom);
        B.u_n[idx] = nk;
        B.u_c[idx] = pz_c(U)[0];
        B.u_r[idx] = __dmul_ru(reduced, RADIUS_SLACK);
    }
    __syncthreads();
}

// ---- one (problem, interval) unit -----------------------------------------------------------------
K1_DI void build_unit(Ctx& c, const Batch& B, int p, int t, double* jrs_mem, int* jrs_n, double* s_unom_r,
                      double* s_dist) {
    const RobotConstants& rc = c_robot;
    const int NJ = B.NJ, T = B.T;
    const double thr = c.thr;
    c.top = 0;
    c.gtop = 0;

    // ---- K2: joint reachable set ----
    double* rot_mem = jrs_mem;
    double* scl_mem = jrs_mem + (MAXJ + 1) * ROT_WORDS;
    if (c.tid < NF) {
        const int i = c.tid;
        jrs_joint(i, t, T, B.q0[size_t(p) * NF + i], B.qd0[size_t(p) * NF + i], B.qdd0[size_t(p) * NF + i], thr,
                  rot_mem + i * ROT_WORDS, &jrs_n[i], scl_mem + (0 * NF + i) * SCL_WORDS, &jrs_n[16 + i],
                  scl_mem + (1 * NF + i) * SCL_WORDS, &jrs_n[24 + i], scl_mem + (2 * NF + i) * SCL_WORDS,
                  &jrs_n[32 + i]);
    } else if (c.tid >= 32 && c.tid < 32 + (NJ + 1 - NF)) {  // fixed joints and the identity after the last one
        const int i = NF + (c.tid - 32);
        double* blk = rot_mem + i * ROT_WORDS;
        for (int e = 0; e < 9; e++) {
            blk[e] = (i < NJ) ? rc.rrpy[i * 9 + e] : ((e % 4 == 0) ? 1.0 : 0.0);
            blk[9 + e] = 0.0;
            blk[18 + e] = 0.0;
        }
        jrs_n[i] = 0;
    }
    __syncthreads();
    Jrs J;
    for (int i = 0; i <= NJ; i++) {
        J.R[i].p = rot_mem + i * ROT_WORDS;
        J.R[i].n = jrs_n[i];
        J.R[i].sz = 9;
    }
    for (int i = 0; i < NF; i++) {
        J.qd[i].p = scl_mem + (0 * NF + i) * SCL_WORDS;
        J.qd[i].n = jrs_n[16 + i];
        J.qd[i].sz = 1;
        J.qda[i].p = scl_mem + (1 * NF + i) * SCL_WORDS;
        J.qda[i].n = jrs_n[24 + i];
        J.qda[i].sz = 1;
        J.qdda[i].p = scl_mem + (2 * NF + i) * SCL_WORDS;
        J.qdda[i].n = jrs_n[32 + i];
        J.qdda[i].sz = 1;
    }

    // ---- forward kinematics of the link volumes (KPR/Dynamics.cu:69-81) ----
    {
        PZH FK_R = pz_alloc(c, 0, 9);
        PZH FK_T = pz_alloc(c, 0, 3);
        if (!c.fail) {
            if (c.tid < 27) FK_R.p[c.tid] = (c.tid < 9 && c.tid % 4 == 0) ? 1.0 : 0.0;
            if (c.tid >= 32 && c.tid < 41) FK_T.p[c.tid - 32] = 0.0;
        }
        __syncthreads();
        for (int i = 0; i < NJ; i++) {
            PZH t1 = op_const_mul<2>(c, &rc.trans[3 * i], 0.0, FK_R);
            PZH FK_T2 = op_add<3>(c, FK_T, t1);
            PZH FK_R2 = op_mul33<3, false>(c, FK_R, J.R[i]);
            // link box zonotope: centre + diag(generators) on the x / y / z generator variables, which
            // reuse the hash slots of qde_0 / qdae_0 / qddae_0 (KPR/Dynamics.cu:51-66)
            PZH box = pz_alloc(c, 3, 3);
            if (!c.fail && c.tid == 0) {
                double rad[3] = {0, 0, 0};
                int n = 0;
                u64 kk[3];
                double gg[3][3];
                for (int jx = 0; jx < 3; jx++) {
                    const double g = rc.link_zonotope_generators[i * 3 + jx];
                    if (sqrt(g * g) <= thr) {
                        rad[jx] = fabs(g);
                    } else {
                        kk[n] = 1ull << (14 + 7 * jx);
                        for (int e = 0; e < 3; e++) gg[n][e] = (e == jx) ? g : 0.0;
                        n++;
                    }
                }
                for (int e = 0; e < 3; e++) {
                    box.p[e] = rc.link_zonotope_center[i * 3 + e];
                    box.p[3 + e] = rad[e];
                    box.p[6 + e] = rad[e];
                }
                u64* pk = reinterpret_cast<u64*>(box.p + 9);
                for (int q = 0; q < n; q++) pk[q] = kk[q];
                for (int q = 0; q < n; q++)
                    for (int e = 0; e < 3; e++) box.p[9 + n + q * 3 + e] = gg[q][e];
                c.cnt[0] = n;
            }
            __syncthreads();
            box.n = c.fail ? 0 : c.cnt[0];
            __syncthreads();
            PZH l1 = op_mul33<1, false>(c, FK_R2, box);
            PZH link = op_add<3>(c, l1, FK_T2);
            export_link(c, link, B, p, t, i);
            PZH* keep[2] = {&FK_T2, &FK_R2};
            arena_keep(c, 0, keep, 2);  // slide over the old FK_R / FK_T
            FK_T = FK_T2;
            FK_R = FK_R2;
        }
    }

    // ---- recursive Newton-Euler, forward pass (KPR/Dynamics.cu:83-155) ----
    c.top = 0;
    PZH Fg[MAXJ], Ng[MAXJ];
    PZH la = pz_zero(c, 3), w = pz_zero(c, 3), wa = pz_zero(c, 3), wd = pz_zero(c, 3);
    if (!c.fail && c.tid == 0) la.p[2] = rc.gravity;
    __syncthreads();
    for (int i = 0; i < NJ; i++) {
        const double* pI = &rc.trans[3 * i];
        const double* cI = &rc.com[3 * i];
        const PZH& Rt = J.R[i];  // used transposed
        // state blocks sit at the bottom of the arena in the order of the keep lists below; each step
        // drops what has just died so the live set stays small
        const int mark = c.top;
        PZH la2;
        {
            PZH t1 = op_cross_const(c, wd, pI, false);
            PZH t2 = op_add<3>(c, la, t1);
            arena_keep1(c, mark, t2);
            PZH t3 = op_cross_const(c, wa, pI, false);
            PZH t4 = op_cross(c, w, t3);
            PZH t5 = op_add<3>(c, t2, t4);
            arena_keep1(c, mark, t5);
            la2 = op_mul33<1, true>(c, Rt, t5);
        }
        PZH w2, wa2, wd3;
        if (rc.axes[i] != 0) {
            const int ax = abs(rc.axes[i]) - 1;
            {
                PZH* k1l[4] = {&w, &wa, &wd, &la2};  // la is dead
                arena_keep(c, 0, k1l, 4);
            }
            PZH w1 = op_mul33<1, true>(c, Rt, w);
            w2 = op_add_one_dim(c, w1, J.qd[i], ax);
            {
                PZH* k2l[4] = {&wa, &wd, &la2, &w2};  // w, w1 are dead
                arena_keep(c, 0, k2l, 4);
            }
            PZH wa1 = op_mul33<1, true>(c, Rt, wa);
            PZH wd1 = op_mul33<1, true>(c, Rt, wd);
            {
                PZH* k3l[4] = {&la2, &w2, &wa1, &wd1};  // wa, wd are dead
                arena_keep(c, 0, k3l, 4);
            }
            const int m2 = c.top;
            PZH zero3 = pz_zero(c, 3);
            PZH tmp = op_add_one_dim(c, zero3, J.qd[i], ax);
            PZH t6 = op_cross(c, wa1, tmp);
            PZH wd2 = op_add<3>(c, wd1, t6);
            arena_keep1(c, m2, wd2);
            wa2 = op_add_one_dim(c, wa1, J.qda[i], ax);
            wd3 = op_add_one_dim(c, wd2, J.qdda[i], ax);
            PZH* k4l[4] = {&la2, &w2, &wa2, &wd3};  // wa1, wd1, wd2 are dead
            arena_keep(c, 0, k4l, 4);
        } else {
            w2 = op_mul33<1, true>(c, Rt, w);
            wa2 = op_mul33<1, true>(c, Rt, wa);
            wd3 = op_mul33<1, true>(c, Rt, wd);
            PZH* k4l[4] = {&la2, &w2, &wa2, &wd3};
            arena_keep(c, 0, k4l, 4);
        }
        {
            const int m3 = c.top;
            PZH t7 = op_cross_const(c, wd3, cI, false);
            PZH t8 = op_add<3>(c, la2, t7);
            arena_keep1(c, m3, t8);
            PZH t9 = op_cross_const(c, wa2, cI, false);
            PZH t10 = op_cross(c, w2, t9);
            PZH t11 = op_add<3>(c, t8, t10);
            arena_keep1(c, m3, t11);
            PZH F = op_const_mul<0>(c, &rc.mass[i], rc.mass_uncertainty, t11);
            Fg[i] = spill_global(c, F);
            c.top = m3;
            PZH t12 = op_const_mul<1>(c, &rc.inertia[i * 9], rc.inertia_uncertainty, wd3);
            PZH t13 = op_const_mul<1>(c, &rc.inertia[i * 9], rc.inertia_uncertainty, w2);
            PZH t14 = op_cross(c, wa2, t13);
            PZH N = op_add<3>(c, t12, t14);
            Ng[i] = spill_global(c, N);
            c.top = m3;
        }
        la = la2;
        w = w2;
        wd = wd3;
        wa = wa2;
    }

    // ---- backward pass (KPR/Dynamics.cu:157-180) ----
    c.top = 0;
    PZH f = pz_zero(c, 3), n = pz_zero(c, 3);
    for (int i = NJ - 1; i >= 0; i--) {
        const PZH& Rn = J.R[i + 1];
        const int mark = c.top;
        PZH a1 = op_mul33<1, false>(c, Rn, n);
        PZH a2 = op_add<3>(c, Ng[i], a1);
        arena_keep1(c, mark, a2);
        PZH a3 = op_cross_const(c, Fg[i], &rc.com[3 * i], true);
        PZH a4 = op_add<3>(c, a2, a3);
        arena_keep1(c, mark, a4);
        PZH a5 = op_mul33<1, false>(c, Rn, f);
        PZH a6 = op_cross_const(c, a5, &rc.trans[3 * (i + 1)], true);
        PZH n2 = op_add<3>(c, a4, a6);
        PZH f2 = op_add<3>(c, a5, Fg[i]);
        PZH* keep[2] = {&n2, &f2};
        arena_keep(c, 0, keep, 2);
        n = n2;
        f = f2;
        if (rc.axes[i] != 0) {
            const int ax = abs(rc.axes[i]) - 1;
            const int m2 = c.top;
            LinSrc s0 = {n, ax, 0, 1.0}, s1 = {J.qdda[i], 0, 0, rc.armature[i]};
            PZH u1 = op_lin2<1>(c, s0, s1);
            LinSrc s2 = {u1, 0, 0, 1.0}, s3 = {J.qd[i], 0, 0, rc.damping[i]};
            PZH u2 = op_lin2<1>(c, s2, s3);
            export_torque(c, u2, B, p, t, i, s_unom_r, s_dist);
            c.top = m2;
        }
    }

    // ---- robust-input radius (KPR/armour_main.cu:172-201) ----
    if (!c.fail && c.tid == 0) {
        double rho = 0.0;
        for (int j = 0; j < NF; j++) rho = __dadd_ru(rho, __dmul_ru(s_dist[j], s_dist[j]));
        const double nrm = __dsqrt_ru(rho);
        for (int j = 0; j < NF; j++) {
            double v = __dadd_ru(__dmul_ru(__dmul_ru(rc.alpha, rc.M_max - rc.M_min), rc.eps), 0.5 * s_dist[j]);
            v = __dadd_ru(v, 0.5 * nrm);
            v = __dadd_ru(v, s_unom_r[j]);
            v = __dadd_ru(v, rc.friction[j]);
            B.torque_radius[size_t(p) * NF * T + size_t(j) * T + t] = __dmul_ru(v, RADIUS_SLACK);
        }
    }
    __syncthreads();
}

// ---- kernel ---------------------------------------------------------------------------------------
#ifndef ARMOUR_EMU
#define K1_SMEM_DECL extern __shared__ __align__(16) unsigned char k1_smem[]
#define K1_SMEM_PTR k1_smem
#else
#define K1_SMEM_DECL unsigned char* k1_smem_emu = reinterpret_cast<unsigned char*>(emu::S().dyn_smem)
#define K1_SMEM_PTR k1_smem_emu
#endif

constexpr int K1_FIXED_BYTES = 64 * 4 + 2 * NW * RED_STRIDE * 8 + 16 * 8 + JRS_WORDS * 8;

__global__ void __launch_bounds__(NT, CTAS_PER_SM) k_reachsets(K1Params P) {
    K1_SMEM_DECL;
    unsigned char* sm = K1_SMEM_PTR;
    int* s_int = reinterpret_cast<int*>(sm);                 // [0..8) cnt, [8] unit, [16..40) jrs counts
    double* s_red = reinterpret_cast<double*>(sm + 64 * 4);
    double* s_red2 = s_red + NW * RED_STRIDE;
    double* s_misc = s_red2 + NW * RED_STRIDE;               // [0..7) u_nom radius, [8..15) disturbance radius
    double* s_jrs = s_misc + 16;
    double* s_arena = s_jrs + JRS_WORDS;
    char* s_tab = reinterpret_cast<char*>(s_arena + P.arena_words);

    Ctx c;
    c.tid = threadIdx.x;
    c.lane = c.tid & 31;
    c.warp = c.tid >> 5;
    c.arena = s_arena;
    c.arena_words = P.arena_words;
    c.top = 0;
    c.top_max = 0;
    c.gscr = P.gscr + size_t(blockIdx.x) * P.gscr_words;
    c.gscr_words = P.fn_words;
    c.garena = c.gscr + P.fn_words;
    c.garena_words = P.gscr_words - P.fn_words;
    c.gtop = 0;
    c.tab_s = s_tab;
    c.tab_s_bytes = P.tab_s_bytes;
    c.tab_g = P.gtab + size_t(blockIdx.x) * P.gtab_bytes;
    c.tab_g_bytes = P.gtab_bytes;
    c.red = s_red;
    c.red2 = s_red2;
    c.cnt = s_int;
    c.thr = c_robot.simplify_threshold;
    c.fail = 0;
    c.n_tab_global = 0;

    for (int i = c.tid; i < P.tab_s_bytes / 8; i += NT) reinterpret_cast<u64*>(s_tab)[i] = 0;
    __syncthreads();

    const int nunits = P.units ? P.nunits : P.B.nprob * P.B.T;
    int nfail = 0, ndone = 0;
    for (;;) {
        if (c.tid == 0) s_int[8] = atomicAdd(P.work, 1);
        __syncthreads();
        int unit = s_int[8];
        __syncthreads();
        if (unit >= nunits) break;
        int p, t;
        if (P.units) {
            unit = P.units[unit];
            p = unit / P.B.T;
            t = unit % P.B.T;
        } else {
            p = unit / P.B.T;
            t = P.B.T - 1 - (unit % P.B.T);  // long intervals first
        }
        c.fail = 0;
        build_unit(c, P.B, p, t, s_jrs, s_int + 16, s_misc, s_misc + 8);
        ndone++;
        if (c.fail) {
            nfail++;
            if (c.tid == 0) atomicMax(&P.B.status[p], c.fail);
            // a failed operation may leave a table half-built: restore the all-zero invariant
            for (int i = c.tid; i < P.tab_s_bytes / 8; i += NT) reinterpret_cast<u64*>(s_tab)[i] = 0;
            for (int i = c.tid; i < P.gtab_bytes / 8; i += NT) reinterpret_cast<u64*>(c.tab_g)[i] = 0;
            __syncthreads();
        }
    }
    if (c.tid == 0 && P.stats) {
        atomicMax(&P.stats[0], c.top_max);
        atomicAdd(&P.stats[1], c.n_tab_global);
        atomicAdd(&P.stats[2], nfail);
        atomicAdd(&P.stats[3], ndone);
    }
}

#ifndef ARMOUR_EMU
// ---- host side: scratch buffers and launch ----------------------------------------------------------
struct K1Scratch {
    int* work = nullptr;
    int* stats = nullptr;
    double* gscr = nullptr;
    char* gtab = nullptr;
    int grid = 0;
    int gscr_words = 0, fn_words = 0, gtab_bytes = 0;
    int arena_words = 0, tab_s_bytes = 0;
    size_t smem_bytes = 0;
    int h_stats[4] = {0, 0, 0, 0};
};

inline void k1_scratch_destroy(K1Scratch* s) {
    if (s->work) cudaFree(s->work);
    if (s->stats) cudaFree(s->stats);
    if (s->gscr) cudaFree(s->gscr);
    if (s->gtab) cudaFree(s->gtab);
    *s = K1Scratch();
}

inline cudaError_t k1_scratch_create(K1Scratch* s, const armour_config& cfg, const RobotConstants&, cudaStream_t st) {
    cudaError_t e;
    int dev = 0, sms = 0, smem_optin = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    // two CTAs per SM: half of the SM's shared memory each (1 KB per CTA is reserved by the system)
    int per_cta = (smem_optin + 1024) / CTAS_PER_SM - 1024;
    per_cta &= ~1023;
    const int dyn = per_cta - K1_FIXED_BYTES;
    s->tab_s_bytes = (dyn * 5 / 8) & ~1023;
    s->arena_words = (dyn - s->tab_s_bytes) / 8;
    s->smem_bytes = size_t(K1_FIXED_BYTES) + size_t(s->arena_words) * 8 + s->tab_s_bytes;
    s->grid = CTAS_PER_SM * sms;
    // per-CTA global scratch: F_i / N_i of one unit, and the overflow hash-table pool
    const int capw = cfg.cap_work_monomials;
    s->fn_words = 2 * MAXJ * (9 + capw * 2);
    s->gscr_words = s->fn_words + 24 * capw;
    s->gtab_bytes = 16 * capw * 8 * 7;  // a cross product table for up to ~10 * capw candidate keys
    if ((e = cudaMalloc(&s->work, sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&s->stats, 4 * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&s->gscr, size_t(s->grid) * s->gscr_words * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&s->gtab, size_t(s->grid) * s->gtab_bytes)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(s->gtab, 0, size_t(s->grid) * s->gtab_bytes, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(s->stats, 0, 4 * sizeof(int), st)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_reachsets, cudaFuncAttributeMaxDynamicSharedMemorySize, int(s->smem_bytes));
}

inline cudaError_t launch_reachsets(const Batch& B, K1Scratch& s, cudaStream_t st, int* nlaunch) {
    *nlaunch = 0;
    if (B.nprob == 0) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaMemsetAsync(s.work, 0, sizeof(int), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(B.status, 0, size_t(B.nprob) * sizeof(int), st)) != cudaSuccess) return e;
    K1Params P;
    P.B = B;
    P.work = s.work;
    P.gscr = s.gscr;
    P.gscr_words = s.gscr_words;
    P.fn_words = s.fn_words;
    P.gtab = s.gtab;
    P.gtab_bytes = s.gtab_bytes;
    P.arena_words = s.arena_words;
    P.tab_s_bytes = s.tab_s_bytes;
    P.stats = s.stats;
    P.units = nullptr;
    P.nunits = 0;
    const int nunits = B.nprob * B.T;
    const int grid = nunits < s.grid ? nunits : s.grid;
    k_reachsets<<<grid, NT, s.smem_bytes, st>>>(P);
    *nlaunch = 1;
    return cudaGetLastError();
}
#endif  // ARMOUR_EMU

}  // namespace k1
#ifndef ARMOUR_EMU
using k1::K1Scratch;
using k1::k1_scratch_create;
using k1::k1_scratch_destroy;
using k1::launch_reachsets;
#endif
}  // namespace armour
