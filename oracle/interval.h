// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned: bit-identical to the reference's own sources (oracle/_ref, tests/test_oracle_pinned.py, tests/golden/ref).
//
// Restatement of the Boost.Interval type the reference uses
// (KPR/Headers.h:30-36: interval<double, policies<save_state<rounded_transc_std<double>>,
// checking_base<double>>>; Boost 1.71 per the reference READMEs — Boost is NOT in
// /root/reference, so this follows the library's published algorithm):
//   * + - * / sqrt are rounded outward with true directed rounding.  GCC does not treat
//     fesetround() as a barrier, so directed rounding is obtained from the round-to-nearest
//     result plus an error-free transformation (TwoSum / FMA residual) that says on which
//     side of the exact result it lies; the result equals the IEEE RD / RU value.
//   * cos/sin: rounded_transc_std evaluates libm cos under a directed rounding mode; glibc's
//     cos installs round-to-nearest internally, so cos_down == cos_up == std::cos.
//   * cos(I): fmod by the interval 2*pi, reflect by the interval pi when the lower bound is
//     past pi, then monotone-piece logic; sin(I) = cos(I - pi/2) with pi/2 an interval.
//   * pow(I, 2) for an interval straddling 0: [0, mul_up(m, m)], m = max(-lo, hi).
// Call sites restated: KPR/Trajectory.cu:97-127, KPR/armour_main.cu:177-190.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>

namespace orc {

namespace rnd {
inline double next_up(double x) { return std::nextafter(x, std::numeric_limits<double>::infinity()); }
inline double next_dn(double x) { return std::nextafter(x, -std::numeric_limits<double>::infinity()); }
inline double two_sum_err(double a, double b, double s) {
    const double bb = s - a;
    return (a - (s - bb)) + (b - bb);
}
inline double add_dn(double a, double b) {
    const double s = a + b;
    return (two_sum_err(a, b, s) < 0) ? next_dn(s) : s;
}
inline double add_up(double a, double b) {
    const double s = a + b;
    return (two_sum_err(a, b, s) > 0) ? next_up(s) : s;
}
inline double sub_dn(double a, double b) { return add_dn(a, -b); }
inline double sub_up(double a, double b) { return add_up(a, -b); }
inline double mul_dn(double a, double b) {
    const double p = a * b;
    return (std::fma(a, b, -p) < 0) ? next_dn(p) : p;
}
inline double mul_up(double a, double b) {
    const double p = a * b;
    return (std::fma(a, b, -p) > 0) ? next_up(p) : p;
}
inline double div_dn(double a, double b) {
    const double q = a / b;
    const double r = std::fma(-q, b, a);  // a - q*b, exact
    const bool too_big = (r != 0) && ((r < 0) != (b < 0));  // exact quotient < q
    return too_big ? next_dn(q) : q;
}
inline double sqrt_dn(double x) {
    const double r = std::sqrt(x);
    return (std::fma(r, r, -x) > 0) ? next_dn(r) : r;
}
inline double sqrt_up(double x) {
    const double r = std::sqrt(x);
    return (std::fma(r, r, -x) < 0) ? next_up(r) : r;
}
}  // namespace rnd

struct Interval {
    double lo = 0, hi = 0;
    Interval() = default;
    Interval(double v) : lo(v), hi(v) {}
    Interval(double l, double h) : lo(l), hi(h) {}
    double lower() const { return lo; }
    double upper() const { return hi; }
};

inline double getCenter(const Interval& a) { return (a.lo + a.hi) * 0.5; }  // KPR/PZsparse.cu:10-12
inline double getRadius(const Interval& a) { return (a.hi - a.lo) * 0.5; }  // KPR/PZsparse.cu:14-16

inline Interval operator-(const Interval& a) { return Interval(-a.hi, -a.lo); }
inline Interval operator+(const Interval& a, const Interval& b) {
    return Interval(rnd::add_dn(a.lo, b.lo), rnd::add_up(a.hi, b.hi));
}
inline Interval operator+(double a, const Interval& b) { return Interval(rnd::add_dn(a, b.lo), rnd::add_up(a, b.hi)); }
inline Interval operator+(const Interval& a, double b) { return b + a; }
inline Interval operator-(const Interval& a, const Interval& b) {
    return Interval(rnd::sub_dn(a.lo, b.hi), rnd::sub_up(a.hi, b.lo));
}
inline Interval operator-(const Interval& a, double b) { return Interval(rnd::sub_dn(a.lo, b), rnd::sub_up(a.hi, b)); }
inline Interval operator*(const Interval& x, const Interval& y) {
    // Boost's sign-case table selects, per case, the same extreme products this min/max finds.
    const double l = std::min(std::min(rnd::mul_dn(x.lo, y.lo), rnd::mul_dn(x.lo, y.hi)),
                              std::min(rnd::mul_dn(x.hi, y.lo), rnd::mul_dn(x.hi, y.hi)));
    const double h = std::max(std::max(rnd::mul_up(x.lo, y.lo), rnd::mul_up(x.lo, y.hi)),
                              std::max(rnd::mul_up(x.hi, y.lo), rnd::mul_up(x.hi, y.hi)));
    return Interval(l, h);
}
inline Interval operator*(double y, const Interval& x) {
    if (y < 0) return Interval(rnd::mul_dn(y, x.hi), rnd::mul_up(y, x.lo));
    if (y == 0) return Interval(0.0, 0.0);
    return Interval(rnd::mul_dn(y, x.lo), rnd::mul_up(y, x.hi));
}
inline Interval operator*(const Interval& x, double y) { return y * x; }

// pow(I, 2) as boost::numeric::pow(interval, int) evaluates it for pwr == 2
inline Interval pow2(const Interval& x) {
    if (x.hi < 0) return Interval(rnd::mul_dn(-x.hi, -x.hi), rnd::mul_up(-x.lo, -x.lo));
    if (x.lo < 0) {
        const double m = std::max(-x.lo, x.hi);
        return Interval(0.0, rnd::mul_up(m, m));
    }
    return Interval(rnd::mul_dn(x.lo, x.lo), rnd::mul_up(x.hi, x.hi));
}

inline Interval sqrt(const Interval& x) {
    const double l = !(x.lo > 0) ? 0.0 : rnd::sqrt_dn(x.lo);
    return Interval(l, rnd::sqrt_up(x.hi));
}

namespace piconst {
// the two doubles bracketing pi (boost::numeric::interval_lib::constants)
constexpr double pi_lower = 3.141592653589793115997963468544185161590576171875;
constexpr double pi_upper = 3.141592653589793560087173318606801331043243408203125;
}  // namespace piconst

inline Interval interval_fmod(const Interval& x, const Interval& y) {
    const double yb = (x.lo < 0) ? y.lo : y.hi;
    const double n = std::floor(rnd::div_dn(x.lo, yb));
    return x - n * y;
}

inline Interval cos(const Interval& x) {
    const Interval pi2(piconst::pi_lower * 2, piconst::pi_upper * 2);
    const Interval pi(piconst::pi_lower, piconst::pi_upper);
    Interval tmp = interval_fmod(x, pi2);
    if (tmp.hi - tmp.lo >= pi2.lo) return Interval(-1.0, 1.0);  // width() is rounded up in Boost; irrelevant here
    if (tmp.lo >= piconst::pi_upper) return -cos(tmp - pi);
    const double l = tmp.lo, u = tmp.hi;
    if (u <= piconst::pi_lower) return Interval(std::cos(u), std::cos(l));
    if (u <= pi2.lo) return Interval(-1.0, std::cos(std::min(rnd::sub_dn(pi2.lo, u), l)));
    return Interval(-1.0, 1.0);
}

inline Interval sin(const Interval& x) {
    const Interval pi_half(piconst::pi_lower * 0.5, piconst::pi_upper * 0.5);
    return cos(x - pi_half);
}

}  // namespace orc
