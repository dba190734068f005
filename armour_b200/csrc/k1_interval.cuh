// Outward-rounded interval arithmetic on the device, for the two places where the reference planner
// uses Boost.Interval (KPR/Headers.h:30-36): the Lagrange remainder of the first-order cos / sin Taylor
// expansion in BezierCurve::makePolyZono (KPR/Trajectory.cu:97-127) and the robust-input norm
// (KPR/armour_main.cu:179-190).  Bounds are produced with the directed-rounding intrinsics
// (__dadd_rd/_ru, __dmul_rd/_ru, __ddiv_rd, __dsqrt_ru), not by switching a rounding mode.
// cos(I) follows the library's published algorithm: reduce modulo the interval 2*pi, reflect by the
// interval pi, then monotone pieces; sin(I) = cos(I - pi/2).
// (included by k1_reachsets.cuh after k1_pz.cuh, once per kernel configuration; no include guard on purpose)

namespace armour {
namespace K1_NS {

struct Itv {
    double lo, hi;
};
K1_DI Itv iv(double l, double h) {
    Itv r;
    r.lo = l;
    r.hi = h;
    return r;
}
K1_DI double iv_mid(const Itv& a) { return (a.lo + a.hi) * 0.5; }  // getCenter, KPR/PZsparse.cu:10-12
K1_DI double iv_rad(const Itv& a) { return (a.hi - a.lo) * 0.5; }  // getRadius, KPR/PZsparse.cu:14-16
K1_DI Itv iv_neg(const Itv& a) { return iv(-a.hi, -a.lo); }
K1_DI Itv iv_add(const Itv& a, const Itv& b) { return iv(__dadd_rd(a.lo, b.lo), __dadd_ru(a.hi, b.hi)); }
K1_DI Itv iv_addd(double a, const Itv& b) { return iv(__dadd_rd(a, b.lo), __dadd_ru(a, b.hi)); }
K1_DI Itv iv_sub(const Itv& a, const Itv& b) { return iv(__dsub_rd(a.lo, b.hi), __dsub_ru(a.hi, b.lo)); }
K1_DI Itv iv_subd(const Itv& a, double b) { return iv(__dsub_rd(a.lo, b), __dsub_ru(a.hi, b)); }
K1_DI Itv iv_mul(const Itv& x, const Itv& y) {
    const double l = fmin(fmin(__dmul_rd(x.lo, y.lo), __dmul_rd(x.lo, y.hi)),
                          fmin(__dmul_rd(x.hi, y.lo), __dmul_rd(x.hi, y.hi)));
    const double h = fmax(fmax(__dmul_ru(x.lo, y.lo), __dmul_ru(x.lo, y.hi)),
                          fmax(__dmul_ru(x.hi, y.lo), __dmul_ru(x.hi, y.hi)));
    return iv(l, h);
}
K1_DI Itv iv_muld(double y, const Itv& x) {
    if (y < 0) return iv(__dmul_rd(y, x.hi), __dmul_ru(y, x.lo));
    if (y == 0) return iv(0.0, 0.0);
    return iv(__dmul_rd(y, x.lo), __dmul_ru(y, x.hi));
}
K1_DI Itv iv_pow2(const Itv& x) {
    if (x.hi < 0) return iv(__dmul_rd(-x.hi, -x.hi), __dmul_ru(-x.lo, -x.lo));
    if (x.lo < 0) {
        const double m = fmax(-x.lo, x.hi);
        return iv(0.0, __dmul_ru(m, m));
    }
    return iv(__dmul_rd(x.lo, x.lo), __dmul_ru(x.hi, x.hi));
}

constexpr double PI_LO = 3.141592653589793115997963468544185161590576171875;
constexpr double PI_HI = 3.141592653589793560087173318606801331043243408203125;

K1_DI Itv iv_fmod(const Itv& x, const Itv& y) {
    const double yb = (x.lo < 0) ? y.lo : y.hi;
    const double n = floor(__ddiv_rd(x.lo, yb));
    return iv_sub(x, iv_muld(n, y));
}
// cos(x) rounded outward.  The device cos() is accurate to 2 ulp, not correctly rounded (the reference evaluates glibc's
// cos, < 1 ulp, under Boost's rounded_transc_std): three ulp outward make the bound both sound and a superset of the
// reference's; what it adds to a remainder radius is ~1e-18.
K1_DI double cos_dn(double x) {
    const double c = cos(x);
    return fmax(-1.0, __dsub_rd(c, __dadd_ru(__dmul_ru(fabs(c), 7e-16), 1e-300)));
}
K1_DI double cos_up(double x) {
    const double c = cos(x);
    return fmin(1.0, __dadd_ru(c, __dadd_ru(__dmul_ru(fabs(c), 7e-16), 1e-300)));
}
K1_DI Itv iv_cos(Itv x) {
    const Itv pi2 = iv(PI_LO * 2, PI_HI * 2);
    const Itv pi = iv(PI_LO, PI_HI);
    bool negate = false;
    for (int it = 0; it < 4; it++) {
        const Itv tmp = iv_fmod(x, pi2);
        if (tmp.hi - tmp.lo >= pi2.lo) return iv(-1.0, 1.0);
        if (tmp.lo >= PI_HI) {
            x = iv_sub(tmp, pi);
            negate = !negate;
            continue;
        }
        const double l = tmp.lo, u = tmp.hi;
        Itv r;
        if (u <= PI_LO)
            r = iv(cos_dn(u), cos_up(l));
        else if (u <= pi2.lo)
            r = iv(-1.0, cos_up(fmin(__dsub_rd(pi2.lo, u), l)));
        else
            r = iv(-1.0, 1.0);
        return negate ? iv_neg(r) : r;
    }
    return iv(-1.0, 1.0);
}
K1_DI Itv iv_sin(const Itv& x) { return iv_cos(iv_sub(x, iv(PI_LO * 0.5, PI_HI * 0.5))); }

}  // namespace K1_NS
}  // namespace armour
