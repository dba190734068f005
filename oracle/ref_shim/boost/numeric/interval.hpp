// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Stand-in for Boost.Interval (the reference pins Boost 1.71 only in prose; Boost is NOT in /root/reference
// nor in this image) restricted to what the reference planner uses (KPR/Headers.h:26-36):
//     boost::numeric::interval<double, policies<save_state<rounded_transc_std<double>>, checking_base<double>>>
// with + - * (interval/interval and double/interval), unary -, +=, cos, sin, pow(I, int), sqrt, lower(), upper(); the
// robust-controller sources (MEX/*.cpp, same interval type: MEX/headers.hpp:19-28) add -=, *=, /, assign().
// It restates the library's published algorithms (boost/numeric/interval/{arith,arith2,transc,utility}.hpp):
// outward rounding for every arithmetic operation, cos by reduction modulo the interval 2*pi and monotone
// pieces, sin(x) = cos(x - pi/2), pow by repeated squaring with directed rounding.
// Unlike oracle/interval.h (error-free transformations), directed rounding here REALLY switches the FPU
// rounding mode (fesetround) like Boost's rounded_arith_std, with volatile operands as the barrier — an
// independent second implementation, so that oracle/_ref pins the oracle.
#pragma once
#include <algorithm>
#include <cfenv>
#include <cmath>

namespace boost {
namespace numeric {
namespace interval_lib {

template <class T> struct rounded_transc_std {};
template <class R> struct save_state {};
template <class T> struct checking_base {};
template <class R, class C> struct policies {};

namespace shim {
inline double up(double (*f)(double, double), double a, double b);
#define SHIM_OP(name, expr)                          \
    inline double name##_up(double a, double b) {    \
        volatile double x = a, y = b;                \
        std::fesetround(FE_UPWARD);                  \
        volatile double r = expr;                    \
        std::fesetround(FE_TONEAREST);               \
        return r;                                    \
    }                                                \
    inline double name##_down(double a, double b) {  \
        volatile double x = a, y = b;                \
        std::fesetround(FE_DOWNWARD);                \
        volatile double r = expr;                    \
        std::fesetround(FE_TONEAREST);               \
        return r;                                    \
    }
SHIM_OP(add, x + y)
SHIM_OP(sub, x - y)
SHIM_OP(mul, x * y)
SHIM_OP(div, x / y)
#undef SHIM_OP
inline double sqrt_up(double a) {
    volatile double x = a;
    std::fesetround(FE_UPWARD);
    volatile double r = std::sqrt(x);
    std::fesetround(FE_TONEAREST);
    return r;
}
inline double sqrt_down(double a) {
    volatile double x = a;
    std::fesetround(FE_DOWNWARD);
    volatile double r = std::sqrt(x);
    std::fesetround(FE_TONEAREST);
    return r;
}
// rounded_transc_std: libm cos called under a directed mode; glibc's cos works in round-to-nearest
// internally whatever the caller's mode, so both directions return std::cos(x)
inline double cos_up(double x) { return std::cos(x); }
inline double cos_down(double x) { return std::cos(x); }
inline double int_down(double x) { return std::floor(x); }
// constants of boost/numeric/interval/constants.hpp (double)
constexpr double pi_lower = 3.141592653589793116, pi_upper = 3.141592653589793560;
constexpr double pi_half_lower = pi_lower / 2, pi_half_upper = pi_upper / 2;
constexpr double pi_twice_lower = pi_lower * 2, pi_twice_upper = pi_upper * 2;
}  // namespace shim
}  // namespace interval_lib

template <class T, class Policies>
class interval {
    T lo_, hi_;

public:
    interval() : lo_(0), hi_(0) {}
    interval(const T& v) : lo_(v), hi_(v) {}
    interval(const T& l, const T& u) : lo_(l), hi_(u) {}
    const T& lower() const { return lo_; }
    const T& upper() const { return hi_; }
    interval& operator+=(const interval& o) { return *this = *this + o; }
    interval& operator+=(const T& o) { return *this = *this + o; }
    interval& operator-=(const interval& o) { return *this = *this - o; }
    interval& operator*=(const interval& o) { return *this = *this * o; }
    interval& operator/=(const interval& o) { return *this = *this / o; }
    void assign(const T& l, const T& u) {  // interval.hpp: assign(); checking_base: l <= u is the caller's duty
        lo_ = l;
        hi_ = u;
    }
};

namespace ish = interval_lib::shim;
#define IV template <class T, class P> inline interval<T, P>
#define I interval<T, P>

IV operator+(const I& x, const I& y) { return I(ish::add_down(x.lower(), y.lower()), ish::add_up(x.upper(), y.upper())); }
IV operator+(const I& x, const T& y) { return I(ish::add_down(x.lower(), y), ish::add_up(x.upper(), y)); }
IV operator+(const T& x, const I& y) { return y + x; }
IV operator-(const I& x) { return I(-x.upper(), -x.lower()); }
IV operator-(const I& x, const I& y) { return I(ish::sub_down(x.lower(), y.upper()), ish::sub_up(x.upper(), y.lower())); }
IV operator-(const I& x, const T& y) { return I(ish::sub_down(x.lower(), y), ish::sub_up(x.upper(), y)); }
IV operator-(const T& x, const I& y) { return I(ish::sub_down(x, y.upper()), ish::sub_up(x, y.lower())); }
// arith.hpp operator*(interval, interval): sign case analysis == outward-rounded min / max of the four products
IV operator*(const I& x, const I& y) {
    const T xl = x.lower(), xu = x.upper(), yl = y.lower(), yu = y.upper();
    if (xl < 0) {
        if (xu > 0) {
            if (yl < 0) {
                if (yu > 0) return I(std::min(ish::mul_down(xl, yu), ish::mul_down(xu, yl)), std::max(ish::mul_up(xl, yl), ish::mul_up(xu, yu)));
                return I(ish::mul_down(xu, yl), ish::mul_up(xl, yl));
            }
            if (yu > 0) return I(ish::mul_down(xl, yu), ish::mul_up(xu, yu));
            return I(T(0), T(0));
        }
        if (yl < 0) {
            if (yu > 0) return I(ish::mul_down(xl, yu), ish::mul_up(xl, yl));
            return I(ish::mul_down(xu, yu), ish::mul_up(xl, yl));
        }
        if (yu > 0) return I(ish::mul_down(xl, yu), ish::mul_up(xu, yl));
        return I(T(0), T(0));
    }
    if (xu > 0) {
        if (yl < 0) {
            if (yu > 0) return I(ish::mul_down(xu, yl), ish::mul_up(xu, yu));
            return I(ish::mul_down(xu, yl), ish::mul_up(xl, yu));
        }
        if (yu > 0) return I(ish::mul_down(xl, yl), ish::mul_up(xu, yu));
        return I(T(0), T(0));
    }
    return I(T(0), T(0));
}
IV operator*(const T& x, const I& y) {
    if (x < 0) return I(ish::mul_down(x, y.upper()), ish::mul_up(x, y.lower()));
    if (x == 0) return I(T(0), T(0));
    return I(ish::mul_down(x, y.lower()), ish::mul_up(x, y.upper()));
}
IV operator*(const I& x, const T& y) { return y * x; }
// arith.hpp operator/(interval, interval) for a divisor that does not contain zero (the only case the reference's
// sources can reach: Vector::normalize() of a twist axis); a divisor containing zero gives the whole line
IV operator/(const I& x, const I& y) {
    const T xl = x.lower(), xu = x.upper(), yl = y.lower(), yu = y.upper();
    if (!(yl > 0) && !(yu < 0)) return I(-HUGE_VAL, HUGE_VAL);
    if (xu < 0) {
        if (yu < 0) return I(ish::div_down(xu, yl), ish::div_up(xl, yu));
        return I(ish::div_down(xl, yl), ish::div_up(xu, yu));
    }
    if (xl < 0) {
        if (yu < 0) return I(ish::div_down(xu, yu), ish::div_up(xl, yu));
        return I(ish::div_down(xl, yl), ish::div_up(xu, yl));
    }
    if (yu < 0) return I(ish::div_down(xu, yu), ish::div_up(xl, yl));
    return I(ish::div_down(xl, yu), ish::div_up(xu, yl));
}

// utility.hpp / arith2.hpp
IV fmod(const I& x, const I& y) {
    const T& yb = (x.lower() < 0) ? y.lower() : y.upper();
    const T n = ish::int_down(ish::div_down(x.lower(), yb));
    return x - n * y;
}
template <class T, class P>
inline T width(const I& x) { return ish::sub_up(x.upper(), x.lower()); }

// transc.hpp
IV cos(const I& x) {
    const I pi2(ish::pi_twice_lower, ish::pi_twice_upper);
    I tmp = fmod(x, pi2);
    if (width(tmp) >= pi2.lower()) return I(T(-1), T(1));
    if (tmp.lower() >= ish::pi_upper) return -cos(tmp - I(ish::pi_lower, ish::pi_upper));
    const T l = tmp.lower(), u = tmp.upper();
    if (u <= ish::pi_lower) return I(ish::cos_down(u), ish::cos_up(l));
    if (u <= pi2.lower()) return I(T(-1), ish::cos_up(std::min(ish::sub_down(pi2.lower(), u), l)));
    return I(T(-1), T(1));
}
IV sin(const I& x) { return cos(x - I(ish::pi_half_lower, ish::pi_half_upper)); }
IV sqrt(const I& x) {
    const T l = !(x.lower() > 0) ? T(0) : ish::sqrt_down(x.lower());
    return I(l, ish::sqrt_up(x.upper()));
}
namespace interval_lib {
namespace shim {
inline double pow_dn(double x, int pwr) {
    double y = (pwr & 1) ? x : 1.0;
    pwr >>= 1;
    while (pwr > 0) {
        x = mul_down(x, x);
        if (pwr & 1) y = mul_down(x, y);
        pwr >>= 1;
    }
    return y;
}
inline double pow_up(double x, int pwr) {
    double y = (pwr & 1) ? x : 1.0;
    pwr >>= 1;
    while (pwr > 0) {
        x = mul_up(x, x);
        if (pwr & 1) y = mul_up(x, y);
        pwr >>= 1;
    }
    return y;
}
}  // namespace shim
}  // namespace interval_lib
IV pow(const I& x, int pwr) {  // arith2.hpp, pwr > 0
    if (x.upper() < 0) {
        const T yl = ish::pow_dn(-x.upper(), pwr), yu = ish::pow_up(-x.lower(), pwr);
        return (pwr & 1) ? I(-yu, -yl) : I(yl, yu);
    }
    if (x.lower() < 0) {
        if (pwr & 1) return I(-ish::pow_up(-x.lower(), pwr), ish::pow_up(x.upper(), pwr));
        return I(T(0), ish::pow_up(std::max(-x.lower(), x.upper()), pwr));
    }
    return I(ish::pow_dn(x.lower(), pwr), ish::pow_up(x.upper(), pwr));
}
#undef IV
#undef I

}  // namespace numeric
}  // namespace boost
