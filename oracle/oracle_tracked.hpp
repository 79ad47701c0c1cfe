// oracle_tracked.hpp — running first-order rounding-error analysis of the oracle.
// TEST INFRASTRUCTURE ONLY (see oracle_base.hpp).
//
// `Tr` is a Float64 value paired with an absolute error bound.  Every arithmetic
// operation and elementary function propagates the incoming bounds to first order
// and adds its own rounding (u = 2^-53 for + - * / sqrt; 1 ulp = 2u for library
// functions).  Instantiating the oracle's templates with FT = Tr evaluates the
// REFERENCE ALGORITHM, in the reference's operation order, and returns for every
// output the bound of the rounding error that algorithm itself can accumulate at
// that point.  The parity tests use it where the pure 1e-12 relative criterion is
// not meaningful: wherever the reference algorithm subtracts nearly equal numbers
// (q_vap - q_sat near saturation, 1 - q_l/(q_l+q_r), aR - bR/(1+cR/λ) ...) any two
// correct Float64 implementations (Julia's own libm vs glibc vs CUDA) differ by an
// amount of this size, and no smaller tolerance can be met by anyone.
#pragma once
#include <cmath>
#include <limits>

namespace orc {

struct Tr {
    double v;  // value (identical to the plain Float64 evaluation)
    double e;  // first-order absolute error bound
    Tr() : v(0), e(0) {}
    Tr(double x) : v(x), e(0) {}  // NOLINT: implicit on purpose (literals, parameters)
    Tr(double x, double err) : v(x), e(err) {}
    explicit operator double() const { return v; }
    explicit operator bool() const = delete;
};

namespace trk {
constexpr double U = 1.1102230246251565e-16;  // unit roundoff 2^-53
constexpr double UF = 2.220446049250313e-16;  // 1 ulp: accuracy of a good libm function
inline double fin(double e, double v) { return (std::isfinite(v) && std::isfinite(e)) ? e : 0.0; }
inline Tr mk(double v, double e) { return Tr(v, fin(e + U * std::fabs(v), v)); }
inline Tr mkf(double v, double e) { return Tr(v, fin(e + UF * std::fabs(v), v)); }
inline double digamma(double x) {
    const double h = 1e-5 * std::fmax(1.0, std::fabs(x));
    return (std::lgamma(x + h) - std::lgamma(x - h)) / (2 * h);
}
}  // namespace trk

inline Tr operator-(Tr a) { return Tr(-a.v, a.e); }
inline Tr operator+(Tr a, Tr b) { return trk::mk(a.v + b.v, a.e + b.e); }
inline Tr operator-(Tr a, Tr b) { return trk::mk(a.v - b.v, a.e + b.e); }
inline Tr operator*(Tr a, Tr b) { return trk::mk(a.v * b.v, std::fabs(a.v) * b.e + std::fabs(b.v) * a.e); }
inline Tr operator/(Tr a, Tr b) {
    const double v = a.v / b.v;
    return trk::mk(v, a.e / std::fabs(b.v) + std::fabs(v) * b.e / std::fabs(b.v));
}
inline Tr& operator+=(Tr& a, Tr b) { a = a + b; return a; }
inline Tr& operator-=(Tr& a, Tr b) { a = a - b; return a; }
inline Tr& operator*=(Tr& a, Tr b) { a = a * b; return a; }
inline Tr& operator/=(Tr& a, Tr b) { a = a / b; return a; }
inline bool operator<(Tr a, Tr b) { return a.v < b.v; }
inline bool operator>(Tr a, Tr b) { return a.v > b.v; }
inline bool operator<=(Tr a, Tr b) { return a.v <= b.v; }
inline bool operator>=(Tr a, Tr b) { return a.v >= b.v; }
inline bool operator==(Tr a, Tr b) { return a.v == b.v; }
inline bool operator!=(Tr a, Tr b) { return a.v != b.v; }

inline Tr exp_(Tr a) { const double v = std::exp(a.v); return trk::mkf(v, std::fabs(v) * a.e); }
inline Tr log_(Tr a) { return trk::mkf(std::log(a.v), a.e / std::fabs(a.v)); }
inline Tr log2_(Tr a) { return trk::mkf(std::log2(a.v), a.e / std::fabs(a.v) / 0.6931471805599453); }
inline Tr log10_(Tr a) { return trk::mkf(std::log10(a.v), a.e / std::fabs(a.v) / 2.302585092994046); }
inline Tr log1p_(Tr a) { return trk::mkf(std::log1p(a.v), a.e / std::fabs(1 + a.v)); }
inline Tr expm1_(Tr a) { return trk::mkf(std::expm1(a.v), std::exp(a.v) * a.e); }
inline Tr sqrt_(Tr a) { const double v = std::sqrt(a.v); return trk::mk(v, v > 0 ? a.e / (2 * v) : 0.0); }
inline Tr cbrt_(Tr a) {
    const double v = std::cbrt(a.v);
    return trk::mkf(v, a.v != 0 ? std::fabs(v) * a.e / (3 * std::fabs(a.v)) : 0.0);
}
inline Tr pow_(Tr a, Tr b) {
    const double v = std::pow(a.v, b.v);
    double e = 0;
    if (a.v != 0) e = std::fabs(v) * (std::fabs(b.v) * a.e / std::fabs(a.v) + std::fabs(std::log(std::fabs(a.v))) * b.e);
    return trk::mkf(v, e);
}
inline Tr tgamma_(Tr a) {
    const double v = std::tgamma(a.v);
    return Tr(v, trk::fin(std::fabs(v) * std::fabs(trk::digamma(a.v)) * a.e + 4 * trk::UF * std::fabs(v), v));
}
inline Tr lgamma_(Tr a) {
    const double v = std::lgamma(a.v);
    return Tr(v, trk::fin(std::fabs(trk::digamma(a.v)) * a.e + 4 * trk::UF * std::fmax(std::fabs(v), 1.0), v));
}
inline Tr erf_(Tr a) { return trk::mkf(std::erf(a.v), 1.1283791670955126 * std::exp(-a.v * a.v) * a.e); }
inline Tr erfc_(Tr a) { return trk::mkf(std::erfc(a.v), 1.1283791670955126 * std::exp(-a.v * a.v) * a.e); }
inline Tr tanh_(Tr a) { const double v = std::tanh(a.v); return trk::mkf(v, (1 - v * v) * a.e); }
inline Tr atanh_(Tr a) { return trk::mkf(std::atanh(a.v), a.e / std::fabs(1 - a.v * a.v)); }
inline Tr fabs_(Tr a) { return Tr(std::fabs(a.v), a.e); }
inline bool isfinite_(Tr a) { return std::isfinite(a.v); }
inline bool isinf_(Tr a) { return std::isinf(a.v); }
inline bool isnan_(Tr a) { return std::isnan(a.v); }

}  // namespace orc
