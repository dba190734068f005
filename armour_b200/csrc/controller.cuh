// K5: interval Newton-Euler pass and robust input of ARMOUR's tracking controller, batched over sampled states
// (SURVEY.md 8f-4).
//
// The reference evaluates, once per control step and for ONE state, passRNEA (nominal model, doubles) and passRNEA_Int
// (model with +-eps mass / inertia uncertainty, Boost intervals) and turns the gap between the two into the robust
// input v (MEX/rnea.cpp:6-187, MEX/spatial.cpp, MEX/spatial_interval.cpp, MEX/robust_controller.cpp:67-181, called
// from MATLAB through MEX/kinova_controller.cpp).  Here one thread carries one sampled state through the same
// recursion, so that a validation sweep over 10^5..10^7 states (ultimate-bound checks, torque-limit margins) is one
// launch.  The arithmetic is the reference's, operation for operation:
//   * intervals round outward at every + - * exactly like boost::numeric::interval with rounded_transc_std
//     (lower bound rounded down, upper bound rounded up; products by the extreme of the four end-point products);
//   * a 3-term inner product is p0 + (p1 + p2): Eigen evaluates a coefficient of a fixed-size product as the sum()
//     of a 3-vector, which it unrolls by halves (ProductEvaluators.h / Redux.h) — for a non-associative scalar the
//     order is visible in the last bit;
//   * sin / cos of a joint angle are doubles (std::sin, std::cos in the reference): the host-pointer entry points
//     compute them with the host's libm, as the reference does, and upload them; the device-pointer entry points use
//     the device's sincos (<= 2 ulp from libm's), which moves an end point by a few ulp.
// The recursion is written once, as a template over the scalar (double: nominal model, Itv: interval model).
#pragma once
#include <cuda_runtime.h>

namespace armour {
namespace ctl {

constexpr int MAXJ = 8;  // joints of a model file (the reference's files: 7)
#define CDI __device__ __forceinline__

struct Itv {
    double lo, hi;
};
CDI Itv itv(double l, double h) {
    Itv r;
    r.lo = l;
    r.hi = h;
    return r;
}
template <class T> CDI T pt(double v);
template <> CDI double pt<double>(double v) { return v; }
template <> CDI Itv pt<Itv>(double v) { return itv(v, v); }

CDI double s_add(double a, double b) { return a + b; }
CDI double s_sub(double a, double b) { return a - b; }
CDI double s_mul(double a, double b) { return a * b; }
CDI double s_neg(double a) { return -a; }
CDI double s_addd(double a, double b) { return a + b; }
CDI Itv s_add(const Itv& a, const Itv& b) { return itv(__dadd_rd(a.lo, b.lo), __dadd_ru(a.hi, b.hi)); }
CDI Itv s_sub(const Itv& a, const Itv& b) { return itv(__dsub_rd(a.lo, b.hi), __dsub_ru(a.hi, b.lo)); }
CDI Itv s_neg(const Itv& a) { return itv(-a.hi, -a.lo); }
CDI Itv s_addd(const Itv& a, double b) { return itv(__dadd_rd(a.lo, b), __dadd_ru(a.hi, b)); }  // interval += double
// plain compare-and-select (fmin / fmax carry NaN handling that costs several moves per call; no NaN can occur here)
CDI double lesser(double a, double b) { return a < b ? a : b; }
CDI double greater(double a, double b) { return a > b ? a : b; }
// interval * interval: the sign cases of boost/numeric/interval/arith.hpp select exactly these extremes
CDI Itv s_mul(const Itv& x, const Itv& y) {
    const double l = lesser(lesser(__dmul_rd(x.lo, y.lo), __dmul_rd(x.lo, y.hi)), lesser(__dmul_rd(x.hi, y.lo), __dmul_rd(x.hi, y.hi)));
    const double h = greater(greater(__dmul_ru(x.lo, y.lo), __dmul_ru(x.lo, y.hi)), greater(__dmul_ru(x.hi, y.lo), __dmul_ru(x.hi, y.hi)));
    return itv(l, h);
}
// scalar (double) times T; for an interval the double acts as the point interval [y, y] (Eigen promotes the scalar to the
// matrix's scalar type; boost's interval * T gives the same end points)
CDI double s_muld(double x, double y) { return x * y; }
CDI Itv s_muld(const Itv& x, double y) {
    if (y < 0) return itv(__dmul_rd(y, x.hi), __dmul_ru(y, x.lo));
    if (y == 0) return itv(0.0, 0.0);
    return itv(__dmul_rd(y, x.lo), __dmul_ru(y, x.hi));
}

template <class T> struct V3 {
    T x[3];
};
template <class T> struct M3 {
    T a[9];  // row major
};
template <class T> CDI T sum3(const T& p0, const T& p1, const T& p2) { return s_add(p0, s_add(p1, p2)); }
template <class T> CDI V3<T> vadd(const V3<T>& a, const V3<T>& b) {
    V3<T> r;
    for (int i = 0; i < 3; i++) r.x[i] = s_add(a.x[i], b.x[i]);
    return r;
}
template <class T> CDI V3<T> vsub(const V3<T>& a, const V3<T>& b) {
    V3<T> r;
    for (int i = 0; i < 3; i++) r.x[i] = s_sub(a.x[i], b.x[i]);
    return r;
}
template <class T> CDI V3<T> vneg(const V3<T>& a) {
    V3<T> r;
    for (int i = 0; i < 3; i++) r.x[i] = s_neg(a.x[i]);
    return r;
}
template <class T> CDI V3<T> vscale(const V3<T>& a, const T& s) {
    V3<T> r;
    for (int i = 0; i < 3; i++) r.x[i] = s_mul(a.x[i], s);
    return r;
}
template <class T> CDI V3<T> vscaled(const V3<T>& a, double s) {
    V3<T> r;
    for (int i = 0; i < 3; i++) r.x[i] = s_muld(a.x[i], s);
    return r;
}
template <class T> CDI V3<T> cross(const V3<T>& a, const V3<T>& b) {
    V3<T> r;
    r.x[0] = s_sub(s_mul(a.x[1], b.x[2]), s_mul(a.x[2], b.x[1]));
    r.x[1] = s_sub(s_mul(a.x[2], b.x[0]), s_mul(a.x[0], b.x[2]));
    r.x[2] = s_sub(s_mul(a.x[0], b.x[1]), s_mul(a.x[1], b.x[0]));
    return r;
}
template <class T> CDI T dot(const V3<T>& a, const V3<T>& b) {
    return sum3(s_mul(a.x[0], b.x[0]), s_mul(a.x[1], b.x[1]), s_mul(a.x[2], b.x[2]));
}
template <class T> CDI V3<T> mv(const M3<T>& A, const V3<T>& b) {  // A b
    V3<T> r;
    for (int i = 0; i < 3; i++) r.x[i] = sum3(s_mul(A.a[3 * i], b.x[0]), s_mul(A.a[3 * i + 1], b.x[1]), s_mul(A.a[3 * i + 2], b.x[2]));
    return r;
}
template <class T> CDI V3<T> mtv(const M3<T>& A, const V3<T>& b) {  // A^T b
    V3<T> r;
    for (int i = 0; i < 3; i++) r.x[i] = sum3(s_mul(A.a[i], b.x[0]), s_mul(A.a[3 + i], b.x[1]), s_mul(A.a[6 + i], b.x[2]));
    return r;
}
template <class T> CDI M3<T> mm(const M3<T>& A, const M3<T>& B) {  // A B
    M3<T> r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            r.a[3 * i + j] = sum3(s_mul(A.a[3 * i], B.a[j]), s_mul(A.a[3 * i + 1], B.a[3 + j]), s_mul(A.a[3 * i + 2], B.a[6 + j]));
    return r;
}
template <class T> CDI M3<T> transpose(const M3<T>& A) {
    M3<T> r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.a[3 * i + j] = A.a[3 * j + i];
    return r;
}
template <class T> CDI M3<T> mneg(const M3<T>& A) {
    M3<T> r;
    for (int i = 0; i < 9; i++) r.a[i] = s_neg(A.a[i]);
    return r;
}
template <class T> CDI M3<T> hat(const V3<T>& w) {  // the w_hat member of a twist (MEX/spatial_interval.cpp:41-43)
    M3<T> r;
    const T z = pt<T>(0.0);
    r.a[0] = z;            r.a[1] = s_neg(w.x[2]); r.a[2] = w.x[1];
    r.a[3] = w.x[2];       r.a[4] = z;             r.a[5] = s_neg(w.x[0]);
    r.a[6] = s_neg(w.x[1]); r.a[7] = w.x[0];       r.a[8] = z;
    return r;
}

template <class T> struct Xf {  // IntTransform / Transform
    M3<T> R;
    V3<T> p;
};
template <class T> struct Tw {  // IntTwist / Twist (w_hat is rebuilt from w where it is needed)
    V3<T> w, v;
};
template <class T> struct Wr {  // IntWrench / Wrench
    V3<T> tau, f;
};
// X1.apply(X2) (spatial_interval.cpp:229-237)
template <class T> CDI Xf<T> xf_apply(const Xf<T>& X, const Xf<T>& x2) {
    Xf<T> r;
    r.R = mm(X.R, x2.R);
    r.p = vadd(x2.p, mtv(x2.R, X.p));
    return r;
}
template <class T> CDI Xf<T> xf_inverse(const Xf<T>& X) {  // :239-244
    Xf<T> r;
    r.R = transpose(X.R);
    r.p = mv(mneg(X.R), X.p);
    return r;
}
template <class T> CDI Tw<T> xf_apply(const Xf<T>& X, const Tw<T>& z) {  // :170-174
    Tw<T> r;
    r.w = mv(X.R, z.w);
    r.v = mv(X.R, vsub(z.v, cross(X.p, z.w)));
    return r;
}
template <class T> CDI Tw<T> xf_invapply(const Xf<T>& X, const Tw<T>& z) {  // :178-182
    Tw<T> r;
    r.w = mtv(X.R, z.w);
    r.v = vadd(mtv(X.R, z.v), cross(X.p, r.w));
    return r;
}
template <class T> CDI Wr<T> xf_invapply(const Xf<T>& X, const Wr<T>& w) {  // :192-196
    Wr<T> r;
    const V3<T> rf = mtv(X.R, w.f);
    r.tau = vadd(mtv(X.R, w.tau), cross(X.p, rf));
    r.f = rf;
    return r;
}
// IntTransform(zeta, theta) with sin(theta), cos(theta) given (:147-157)
template <class T> CDI Xf<T> xf_joint(const Tw<T>& zeta, double s, double c) {
    const M3<T> wh = hat(zeta.w);
    M3<T> a, b;
    const double c1 = 1 - c;
    for (int i = 0; i < 9; i++) {
        a.a[i] = s_muld(wh.a[i], s);   // w_hat * sin
        b.a[i] = s_muld(wh.a[i], c1);  // (1 - cos) * w_hat
    }
    const M3<T> bb = mm(b, wh);
    Xf<T> X;
    M3<T> imr;  // identity - R
    for (int i = 0; i < 9; i++) {
        const T id = pt<T>((i % 4 == 0) ? 1.0 : 0.0);
        X.R.a[i] = s_add(s_add(id, a.a[i]), bb.a[i]);
        imr.a[i] = s_sub(id, X.R.a[i]);
    }
    const V3<T> p0 = mv(mm(imr, wh), zeta.v);
    X.p = mv(mneg(transpose(X.R)), p0);
    return X;
}
template <class T> CDI Tw<T> tw_cross(const Tw<T>& a, const Tw<T>& z2) {  // IntTwist::cross(IntTwist) :79-83
    Tw<T> r;
    const M3<T> wh = hat(a.w);
    r.w = mv(wh, z2.w);
    r.v = vadd(mv(wh, z2.v), cross(a.v, z2.w));
    return r;
}
template <class T> CDI Tw<T> tw_add(const Tw<T>& a, const Tw<T>& b) {
    Tw<T> r;
    r.w = vadd(a.w, b.w);
    r.v = vadd(a.v, b.v);
    return r;
}
template <class T> CDI Tw<T> tw_scaled(const Tw<T>& a, double s) {  // Sb * qd: the double becomes the scalar type first
    Tw<T> r;
    r.w = vscaled(a.w, s);
    r.v = vscaled(a.v, s);
    return r;
}

// one joint of the model, for scalar T; the second block is derived from the first once per model (k_model_setup):
// it does not depend on the state, the reference recomputes it in every call (rnea.cpp:126-146)
template <class T> struct JointModel {
    Tw<T> S;        // joint twist
    Xf<T> X;        // XTree
    T m;            // inertia
    M3<T> Ibar, mch;
    T transI;
    Xf<T> Xbw;      // body to world
    Tw<T> Sb;       // screw axis in the body frame
    Xf<T> Xinv;     // XTree.inverse()
};
struct Model {
    int nj;
    int chain;           // 1: parent[i] == i - 1 for every joint
    int parent[MAXJ];
    double gravity[3];   // linear part of the gravity twist
    double friction[MAXJ], damping[MAXJ];
    JointModel<double> nom[MAXJ];
    JointModel<Itv> iv[MAXJ];
};
template <class T> CDI const JointModel<T>* joints(const Model* M);
template <> CDI const JointModel<double>* joints<double>(const Model* M) { return M->nom; }
template <> CDI const JointModel<Itv>* joints<Itv>(const Model* M) { return M->iv; }

template <class T> __device__ void model_setup(JointModel<T>* J, const int* parent, int nj) {
    for (int i = 0; i < nj; i++) {
        const int li = parent[i];
        J[i].Xbw = (li != -1) ? xf_apply(J[li].Xbw, J[i].X) : J[i].X;
        J[i].Sb = xf_invapply(J[i].Xbw, J[i].S);
        J[i].Xinv = xf_inverse(J[i].X);
    }
}
__global__ void k_model_setup(Model* M) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        model_setup<double>(M->nom, M->parent, M->nj);
        model_setup<Itv>(M->iv, M->parent, M->nj);
    }
}

// The joint transforms of one state, Xli[i] = IntTransform(Sb[i], -q[i]).apply(XTree[i].inverse()) (MEX/rnea.cpp:139-146):
// they depend on the angles only, so the controller update computes them once for its two interval passes.
template <class T>
__device__ void rnea_transforms(const Model* M, const double* sn, const double* cs, Xf<T>* Xli) {
    const JointModel<T>* J = joints<T>(M);
    for (int i = 0; i < M->nj; i++) Xli[i] = xf_apply(xf_joint(J[i].Sb, sn[i], cs[i]), J[i].Xinv);
}

// passRNEA / passRNEA_Int (MEX/rnea.cpp:6-94 / 96-187) for one state, given its joint transforms.
// Requires parent[i] < i (a spanning tree numbered from the base, as the reference's files are).
// ZERO_VEL: the call with qd = qda = 0 (M(q) r of the controller, robust_controller.cpp:148): every velocity twist is the zero
// interval, so the velocity products are [0, 0] and adding them changes nothing (x + 0 is exact in every rounding direction);
// they are left out, the result is the reference's.
// CHAIN: parent[i] == i - 1 for every joint (a serial arm, like the reference's models): the twists of the parent are the ones
// just computed, so one slot per twist is enough instead of one per joint (2.3 KB less local memory per interval pass).
template <class T, bool ZERO_VEL, bool CHAIN>
__device__ void rnea_pass_impl(const Model* M, const Xf<T>* Xli, const double* qd, const double* qda, const double* qdd,
                               bool friction, bool gravity, T* tau) {
    const JointModel<T>* J = joints<T>(M);
    const int nj = M->nj;
    Wr<T> f[MAXJ];
    constexpr int SLOTS = CHAIN ? 1 : MAXJ;
    Tw<T> v_[SLOTS], va_[SLOTS], a_[SLOTS];
#define v(i) v_[CHAIN ? 0 : (i)]
#define va(i) va_[CHAIN ? 0 : (i)]
#define a(i) a_[CHAIN ? 0 : (i)]
    Tw<T> neg_g;
    for (int k = 0; k < 3; k++) {
        neg_g.w.x[k] = pt<T>(0.0);
        neg_g.v.x[k] = gravity ? s_neg(pt<T>(M->gravity[k])) : pt<T>(0.0);
    }
    for (int i = 0; i < nj; i++) {
        const int li = M->parent[i];
        const Tw<T> Sb = J[i].Sb;
        const Tw<T> sqdd = tw_scaled(Sb, qdd[i]);
        if (ZERO_VEL) {
            a(i) = li == -1 ? tw_add(xf_apply(Xli[i], neg_g), sqdd) : tw_add(xf_apply(Xli[i], a(li)), sqdd);
            f[i].tau = vadd(mv(J[i].Ibar, a(i).w), mv(J[i].mch, a(i).v));  // I.apply(a)
            f[i].f = vsub(vscale(a(i).v, J[i].m), mv(J[i].mch, a(i).w));
            continue;
        }
        const Tw<T> sqd = tw_scaled(Sb, qd[i]), sqda = tw_scaled(Sb, qda[i]);
        if (li == -1) {
            v(i) = sqd;
            va(i) = sqda;
            a(i) = tw_add(tw_add(xf_apply(Xli[i], neg_g), sqdd), tw_cross(v(i), va(i)));
        } else {
            v(i) = tw_add(xf_apply(Xli[i], v(li)), sqd);
            va(i) = tw_add(xf_apply(Xli[i], va(li)), sqda);
            a(i) = tw_add(tw_add(xf_apply(Xli[i], a(li)), sqdd), tw_cross(v(i), sqda));
        }
        // v x I v, the passivity-based way (rnea.cpp:156-160)
        Wr<T> vIv;
        vIv.tau = cross(va(i).w, mv(J[i].Ibar, v(i).w));
        vIv.tau = vadd(vIv.tau, mv(J[i].Ibar, cross(va(i).w, v(i).w)));
        vIv.f = vscale(cross(va(i).w, v(i).v), J[i].m);
        // I.apply(a) (spatial_interval.cpp:131-135)
        Wr<T> Ia;
        Ia.tau = vadd(mv(J[i].Ibar, a(i).w), mv(J[i].mch, a(i).v));
        Ia.f = vsub(vscale(a(i).v, J[i].m), mv(J[i].mch, a(i).w));
        f[i].tau = vadd(Ia.tau, vIv.tau);
        f[i].f = vadd(Ia.f, vIv.f);
    }
    for (int i = nj - 1; i >= 0; i--) {
        const Tw<T> Sb = J[i].Sb;
        T t = s_add(dot(Sb.w, f[i].tau), dot(Sb.v, f[i].f));
        t = s_add(t, s_muld(J[i].transI, qdd[i]));   // transmission inertia
        t = s_addd(t, M->damping[i] * qd[i]);          // damping
        if (friction) t = s_addd(t, M->friction[i] * double((qd[i] > 0) - (qd[i] < 0)));
        tau[i] = t;
        const int li = M->parent[i];
        if (li != -1) {
            const Wr<T> up = xf_invapply(Xli[i], f[i]);
            f[li].tau = vadd(f[li].tau, up.tau);
            f[li].f = vadd(f[li].f, up.f);
        }
    }
}
#undef v
#undef va
#undef a
template <class T, bool ZERO_VEL = false>
__device__ void rnea_pass(const Model* M, const Xf<T>* Xli, const double* qd, const double* qda, const double* qdd, bool friction,
                          bool gravity, T* tau) {
    if (M->chain)
        rnea_pass_impl<T, ZERO_VEL, true>(M, Xli, qd, qda, qdd, friction, gravity, tau);
    else
        rnea_pass_impl<T, ZERO_VEL, false>(M, Xli, qd, qda, qdd, friction, gravity, tau);
}
template <class T>
__device__ void rnea(const Model* M, const double* qd, const double* qda, const double* qdd, const double* sn, const double* cs,
                     bool friction, bool gravity, T* tau) {
    Xf<T> Xli[MAXJ];
    rnea_transforms<T>(M, sn, cs, Xli);
    rnea_pass<T>(M, Xli, qd, qda, qdd, friction, gravity, tau);
}

struct TrigSrc {
    const double* host_sincos;  // [n][nj][2] = sin(-q), cos(-q) from the host's libm, or nullptr: computed here
};
CDI void load_trig(const TrigSrc& t, const double* q, size_t s, int nj, double* sn, double* cs) {
    for (int i = 0; i < nj; i++) {
        if (t.host_sincos) {
            sn[i] = t.host_sincos[(s * nj + i) * 2];
            cs[i] = t.host_sincos[(s * nj + i) * 2 + 1];
        } else {
            sincos(-q[s * nj + i], &sn[i], &cs[i]);
        }
    }
}

constexpr int CTL_THREADS = 128;
#ifndef CTL_MINB
#define CTL_MINB 3  // 168 registers: measured 4.6 ms per 2^20 interval passes (1: 5.6 ms, 4: 4.9 ms, 6: 8.7 ms)
#endif

// tau_lo / tau_hi [n][nj] (interval model) and / or tau [n][nj] (nominal model); either output may be nullptr
__global__ void __launch_bounds__(CTL_THREADS, CTL_MINB)
k_rnea(const Model* __restrict__ M, int n, const double* __restrict__ q, const double* __restrict__ qd, const double* __restrict__ qda,
       const double* __restrict__ qdd, TrigSrc trig, int friction, int gravity, double* __restrict__ tau, double* __restrict__ tau_lo,
       double* __restrict__ tau_hi) {
    const size_t s = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= size_t(n)) return;
    const int nj = M->nj;
    double sn[MAXJ], cs[MAXJ];
    load_trig(trig, q, s, nj, sn, cs);
    if (tau) {
        double t[MAXJ];
        rnea<double>(M, qd + s * nj, qda + s * nj, qdd + s * nj, sn, cs, friction != 0, gravity != 0, t);
        for (int i = 0; i < nj; i++) tau[s * nj + i] = t[i];
    }
    if (tau_lo) {
        Itv t[MAXJ];
        rnea<Itv>(M, qd + s * nj, qda + s * nj, qdd + s * nj, sn, cs, friction != 0, gravity != 0, t);
        for (int i = 0; i < nj; i++) {
            tau_lo[s * nj + i] = t[i].lo;
            tau_hi[s * nj + i] = t[i].hi;
        }
    }
}

struct ControllerGains {
    double Kr[MAXJ];   // diagonal of Kr (MEX/kinova_controller.cpp:36-40)
    double alpha, V_max, r_norm_threshold;
    int friction;      // RobustController::applyFriction (the MEX entry sets it to false)
};
// norm of an nj-vector of doubles in the order of Eigen's SSE2 reduction of a dynamic vector (Redux.h: two packet
// accumulators of two doubles, the remaining packet, the lanes, the scalar tail)
CDI double norm_dyn(const double* x, int n) {
    const int aligned2 = (n / 4) * 4, aligned = (n / 2) * 2;
    double res;
    if (aligned == 0) {
        res = x[0] * x[0];
        for (int i = 1; i < n; i++) res += x[i] * x[i];
        return sqrt(res);
    }
    double a0 = x[0] * x[0], a1 = x[1] * x[1];
    if (aligned > 2) {
        double b0 = x[2] * x[2], b1 = x[3] * x[3];
        for (int i = 4; i < aligned2; i += 4) {
            a0 += x[i] * x[i];
            a1 += x[i + 1] * x[i + 1];
            b0 += x[i + 2] * x[i + 2];
            b1 += x[i + 3] * x[i + 3];
        }
        a0 += b0;
        a1 += b1;
        if (aligned > aligned2) {
            a0 += x[aligned2] * x[aligned2];
            a1 += x[aligned2 + 1] * x[aligned2 + 1];
        }
    }
    res = a0 + a1;
    for (int i = aligned; i < n; i++) res += x[i] * x[i];
    return sqrt(res);
}
CDI double wrap_pi(double a) {  // clamp() of MEX/robust_controller.hpp:11-16
    const double PI = 3.14159265358979323846, TWOPI = 6.283185307179586476925286766559;
    while (a >= PI) a -= TWOPI;
    while (a < -PI) a += TWOPI;
    return a;
}

// RobustController::update, ARMOUR method (MEX/robust_controller.cpp:67-181): u, u_nominal, v [n][nj]; status[n] = 1 where
// the nominal torque falls outside the interval torque (the reference throws there), else 0
__global__ void __launch_bounds__(CTL_THREADS, CTL_MINB)
k_controller_update(const Model* __restrict__ M, int n, ControllerGains G, const double* __restrict__ q, const double* __restrict__ qd,
                    const double* __restrict__ q_des, const double* __restrict__ qd_des, const double* __restrict__ qdd_des, TrigSrc trig,
                    double* __restrict__ u, double* __restrict__ u_nominal, double* __restrict__ v_out, int* __restrict__ status) {
    const size_t s = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= size_t(n)) return;
    const int nj = M->nj;
    double sn[MAXJ], cs[MAXJ], qa_d[MAXJ], qa_dd[MAXJ], r[MAXJ], qdv[MAXJ], zero[MAXJ];
    load_trig(trig, q, s, nj, sn, cs);
    for (int i = 0; i < nj; i++) {
        const double qdiff = wrap_pi(q_des[s * nj + i] - q[s * nj + i]);
        const double ed = qd_des[s * nj + i] - qd[s * nj + i];
        qdv[i] = qd[s * nj + i];
        qa_d[i] = qd_des[s * nj + i] + G.Kr[i] * qdiff;
        qa_dd[i] = qdd_des[s * nj + i] + G.Kr[i] * ed;
        r[i] = ed + G.Kr[i] * qdiff;
        zero[i] = 0.0;
    }
    double un[MAXJ];
    Itv ui[MAXJ];
    rnea<double>(M, qdv, qa_d, qa_dd, sn, cs, G.friction != 0, true, un);
    Xf<Itv> Xli[MAXJ];  // shared by the interval torque pass and the interval M(q) r pass below
    rnea_transforms<Itv>(M, sn, cs, Xli);
    rnea_pass<Itv>(M, Xli, qdv, qa_d, qa_dd, G.friction != 0, true, ui);
    int st = 0;
    double bound[MAXJ];
    for (int i = 0; i < nj; i++) {
        if (un[i] > ui[i].hi || un[i] < ui[i].lo) st = 1;
        const Itv phi = s_sub(ui[i], pt<Itv>(un[i]));
        bound[i] = fmax(fabs(phi.lo), fabs(phi.hi));
    }
    double v[MAXJ];
    for (int i = 0; i < nj; i++) v[i] = 0.0;
    const double r_norm = norm_dyn(r, nj);
    if (r_norm > G.r_norm_threshold) {
        Itv Mr[MAXJ];
        rnea_pass<Itv, true>(M, Xli, zero, zero, r, false, false, Mr);  // M(q) r
        Itv V = pt<Itv>(0.0);
        for (int i = 0; i < nj; i++) V = s_add(V, s_muld(Mr[i], 0.5 * r[i]));
        const double h = -V.hi + G.V_max;
        const double lambda = fmax(0.0, -G.alpha * h / r_norm + norm_dyn(bound, nj));
        for (int i = 0; i < nj; i++) v[i] = -lambda * r[i] / r_norm;
    }
    for (int i = 0; i < nj; i++) {
        u[s * nj + i] = un[i] - v[i];
        if (u_nominal) u_nominal[s * nj + i] = un[i];
        if (v_out) v_out[s * nj + i] = v[i];
    }
    if (status) status[s] = st;
}

}  // namespace ctl
}  // namespace armour
