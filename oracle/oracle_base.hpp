// oracle_base.hpp — CPU restatement of the reference's shared numerics.
//
// TEST INFRASTRUCTURE ONLY.  This directory is the parity oracle for the CUDA
// path: a plain C++ restatement of CliMA/CloudMicrophysics.jl's pointwise
// algorithms, written operation for operation in the order of the Julia source
// (compiled with -ffp-contract=off; the reference only fuses where it writes
// `muladd`).  Nothing in the product (cloudmicrophysics.jl_b200/, libcumicro.so)
// may include, link or call it; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs do.
//
// Third-party arithmetic that is NOT under /root/reference (no Manifest.toml;
// compat bounds in Project.toml:27-42) is restated from the published formulas:
//   Thermodynamics.jl (0.15.4 / 1)   -> struct Thermo below
//   LogExpFunctions.jl (0.3.29 / 1)  -> log1pexp / log1mexp / cloglog below
//   SpecialFunctions.jl (2.7.1)      -> libm tgamma / lgamma / erf / erfc
// Parity pins: tests/test_oracle_goldens.py checks every function here against
// the known-answer literals of the reference's own tests.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <algorithm>

#include "../include/cumicro.h"
#include "oracle_tracked.hpp"

namespace orc {

// ---- elementary functions: libm for float/double/long double, error-propagating
// overloads for Tr (oracle_tracked.hpp).  All oracle code calls these unqualified.
#define ORC_FN1(name) \
    inline float name##_(float x) { return std::name(x); } \
    inline double name##_(double x) { return std::name(x); }
ORC_FN1(exp) ORC_FN1(log) ORC_FN1(log2) ORC_FN1(log10) ORC_FN1(log1p) ORC_FN1(expm1) ORC_FN1(sqrt) ORC_FN1(cbrt)
ORC_FN1(tgamma) ORC_FN1(lgamma) ORC_FN1(erf) ORC_FN1(erfc) ORC_FN1(tanh) ORC_FN1(atanh) ORC_FN1(fabs)
#undef ORC_FN1
inline float pow_(float x, float y) { return std::pow(x, y); }
inline double pow_(double x, double y) { return std::pow(x, y); }
inline bool isfinite_(double x) { return std::isfinite(x); }
inline bool isfinite_(float x) { return std::isfinite(x); }
inline bool isinf_(double x) { return std::isinf(x); }
inline bool isinf_(float x) { return std::isinf(x); }
inline bool isnan_(double x) { return std::isnan(x); }
inline bool isnan_(float x) { return std::isnan(x); }

template <class FT> struct PT;  // parameter-type selector
template <> struct PT<double> {
    using thermo = cumicro_thermo_f64;
    using air = cumicro_air_f64;
    using sb_pdf_c = cumicro_sb_pdf_c_f64;
    using sb_pdf_r = cumicro_sb_pdf_r_f64;
    using sb_acnv = cumicro_sb_acnv_f64;
    using sb_accr = cumicro_sb_accr_f64;
    using sb_self = cumicro_sb_self_f64;
    using sb_brek = cumicro_sb_brek_f64;
    using sb_evap = cumicro_sb_evap_f64;
    using sb2006 = cumicro_sb2006_f64;
    using vel_sb2006 = cumicro_vel_sb2006_f64;
    using vel_stokes = cumicro_vel_stokes_f64;
    using vel_chen_rain = cumicro_vel_chen_rain_f64;
    using vel_chen_small_ice = cumicro_vel_chen_small_ice_f64;
    using vel_chen_large_ice = cumicro_vel_chen_large_ice_f64;
    using params_2m_warm = cumicro_params_2m_warm_f64;
};
template <> struct PT<Tr> : PT<double> {};  // parameters stay plain Float64
template <> struct PT<float> {
    using thermo = cumicro_thermo_f32;
    using air = cumicro_air_f32;
    using sb_pdf_c = cumicro_sb_pdf_c_f32;
    using sb_pdf_r = cumicro_sb_pdf_r_f32;
    using sb_acnv = cumicro_sb_acnv_f32;
    using sb_accr = cumicro_sb_accr_f32;
    using sb_self = cumicro_sb_self_f32;
    using sb_brek = cumicro_sb_brek_f32;
    using sb_evap = cumicro_sb_evap_f32;
    using sb2006 = cumicro_sb2006_f32;
    using vel_sb2006 = cumicro_vel_sb2006_f32;
    using vel_stokes = cumicro_vel_stokes_f32;
    using vel_chen_rain = cumicro_vel_chen_rain_f32;
    using vel_chen_small_ice = cumicro_vel_chen_small_ice_f32;
    using vel_chen_large_ice = cumicro_vel_chen_large_ice_f32;
    using params_2m_warm = cumicro_params_2m_warm_f32;
};

// ---- Julia Base semantics -------------------------------------------------
template <class FT> inline FT jmax(FT a, FT b) { return (a < b) ? b : a; }
template <class FT> inline FT jmin(FT a, FT b) { return (b < a) ? b : a; }
// Base.clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x))
template <class FT> inline FT jclamp(FT x, FT lo, FT hi) { return (x > hi) ? hi : ((x < lo) ? lo : x); }
template <class FT> inline FT ifelse(bool c, FT a, FT b) { return c ? a : b; }
inline Tr jmax(Tr a, Tr b) { return (a < b) ? b : a; }
inline Tr jmin(Tr a, Tr b) { return (b < a) ? b : a; }
inline Tr jclamp(Tr x, Tr lo, Tr hi) { return (x > hi) ? hi : ((x < lo) ? lo : x); }
inline Tr ifelse(bool c, Tr a, Tr b) { return c ? a : b; }
// Threshold mode: when set, the Float64 (and Tr) instantiations use the FLOAT32 method's
// thresholds eps(Float32), ϵ_numerics(Float32).  This evaluates "the Float32 method in exact
// arithmetic" — the true value a Float32 implementation is judged against (its regime
// selection is the Float32 reference's, its arithmetic error-free to ~1e-16).
inline bool& f32_thresholds() {
    static bool flag = false;
    return flag;
}
template <class FT> inline FT eps() { return std::numeric_limits<FT>::epsilon(); }
template <class FT> inline FT inf() { return std::numeric_limits<FT>::infinity(); }
template <> inline double eps<double>() {
    return f32_thresholds() ? double(std::numeric_limits<float>::epsilon()) : std::numeric_limits<double>::epsilon();
}
template <> inline Tr eps<Tr>() { return Tr(eps<double>()); }
template <> inline Tr inf<Tr>() { return Tr(std::numeric_limits<double>::infinity()); }
template <class FT> inline FT pi() { return FT(3.141592653589793238462643383279502884L); }

// ---- Utilities.jl -----------------------------------------------------------
// UT.ϵ_numerics(FT) = cbrt(floatmin(FT))                         UT:318
template <class FT> inline FT eps_numerics() { return cbrt_(std::numeric_limits<FT>::min()); }
template <> inline double eps_numerics<double>() {
    return f32_thresholds() ? double(std::cbrt(std::numeric_limits<float>::min())) : std::cbrt(std::numeric_limits<double>::min());
}
template <> inline Tr eps_numerics<Tr>() { return Tr(eps_numerics<double>()); }
// UT.ϵ_numerics_2M_M / _2M_N / _P3_B = eps(FT)                   UT:325,332,340
template <class FT> inline FT eps_2M() { return eps<FT>(); }
// UT.clamp_to_nonneg                                             UT:296
template <class FT> inline FT clamp_to_nonneg(FT x) { return jmax(FT(0), x); }

// ---- LogExpFunctions.jl (external; restated from its definition) -----------
// log1pexp(x): Maechler (2012) branches with the package's thresholds.
template <class FT> struct Log1pExpCut { static constexpr double a = -36.7368005696771, b = 18.021826694558577, c = 33.23111882352963; };
template <> struct Log1pExpCut<float> { static constexpr double a = -15.942385, b = 9.011913, c = 16.635532; };
template <class FT> inline FT log1pexp(FT x) {
    using C = Log1pExpCut<FT>;
    if (x < FT(C::a)) return exp_(x);
    if (x < FT(C::b)) return log1p_(exp_(x));
    if (x < FT(C::c)) return x + exp_(-x);
    return x;
}
// log1mexp(x) = x < log(1/2) ? log1p(-exp(x)) : log(-expm1(x))
template <class FT> inline FT log1mexp(FT x) {
    const FT loghalf = FT(-0.6931471805599453094172321214581765680755L);
    return (x < loghalf) ? log1p_(-exp_(x)) : log_(-expm1_(x));
}
// cloglog(x) = log(-log1p(-x))
template <class FT> inline FT cloglog(FT x) { return log_(-log1p_(-x)); }

// ---- Thermodynamics.jl (external) --------------------------
template <class FT> struct Thermo {
    const typename PT<FT>::thermo& p;
    explicit Thermo(const typename PT<FT>::thermo& p_) : p(p_) {}
    FT T_freeze() const { return p.T_freeze; }
    FT R_v() const { return p.R_v; }
    FT cv_l() const { return p.cp_l; }
    FT L_v(FT T) const { return p.LH_v0 + (p.cp_v - p.cp_l) * (T - p.T_0); }
    FT L_s(FT T) const { return p.LH_s0 + (p.cp_v - p.cp_i) * (T - p.T_0); }
    FT L_f(FT T) const { return (p.LH_s0 - p.LH_v0) + (p.cp_l - p.cp_i) * (T - p.T_0); }
    FT cp_m(FT qt, FT ql, FT qi) const {
        return p.cp_d + (p.cp_v - p.cp_d) * qt + (p.cp_l - p.cp_v) * ql + (p.cp_i - p.cp_v) * qi;
    }
    FT R_m(FT qt, FT ql, FT qi) const {
        const FT Rv_over_Rd = p.R_v / p.R_d;
        return p.R_d * (1 + (Rv_over_Rd - 1) * qt - Rv_over_Rd * (ql + qi));
    }
    FT p_sat_calc(FT T, FT LH_0, FT dcp) const {
        return p.press_triple * pow_(T / p.T_triple, dcp / p.R_v) *
               exp_((LH_0 - dcp * p.T_0) / p.R_v * (1 / p.T_triple - 1 / T));
    }
    FT p_sat_liq(FT T) const { return p_sat_calc(T, p.LH_v0, p.cp_v - p.cp_l); }
    FT p_sat_ice(FT T) const { return p_sat_calc(T, p.LH_s0, p.cp_v - p.cp_i); }
    // q_vap_from_p_vap(T, ρ, p_v) = p_v / (R_v ρ T)   (TDI:64 p2q)
    FT p2q(FT T, FT rho, FT pv) const { return pv / (p.R_v * rho * T); }
    FT q2p(FT T, FT rho, FT qv) const { return qv * rho * p.R_v * T; }  // TDI:67
    FT q_sat_liq(FT T, FT rho) const { return p2q(T, rho, p_sat_liq(T)); }
    FT q_sat_ice(FT T, FT rho) const { return p2q(T, rho, p_sat_ice(T)); }
    // TDI.q_vap                                                 TDI:60-61
    static FT q_vap(FT qt, FT ql, FT qi) { return clamp_to_nonneg(qt - ql - qi); }
    // TDI.supersaturation_over_{liquid,ice}                      TDI:118-125
    FT supersat_liq(FT qt, FT ql, FT qi, FT rho, FT T) const {
        FT qv = q_vap(qt, ql, qi);
        return qv * (rho * p.R_v * T) / p_sat_liq(T) - 1;
    }
    FT supersat_ice(FT qt, FT ql, FT qi, FT rho, FT T) const {
        FT qv = q_vap(qt, ql, qi);
        return qv * (rho * p.R_v * T) / p_sat_ice(T) - 1;
    }
};

// ---- Common.jl ---------------------------------------------------------------
// CO.G_func_liquid / G_func_ice                                  CO:47-102
template <class FT>
inline FT G_func(const typename PT<FT>::air& aps, FT R_v, FT L, FT p_vs, FT T) {
    const FT e = eps_numerics<FT>();
    FT p_vs_safe = jmax(p_vs, e);
    FT D_vapor_safe = jmax(aps.D_vapor, e);
    FT K_therm_safe = jmax(aps.K_therm, e);
    return 1 / (L / K_therm_safe / T * (L / R_v / T - 1) + R_v * T / D_vapor_safe / p_vs_safe);
}
template <class FT>
inline FT G_func_liquid(const typename PT<FT>::air& aps, const Thermo<FT>& tps, FT T) {
    return G_func<FT>(aps, tps.R_v(), tps.L_v(T), tps.p_sat_liq(T), T);
}
template <class FT>
inline FT G_func_ice(const typename PT<FT>::air& aps, const Thermo<FT>& tps, FT T) {
    return G_func<FT>(aps, tps.R_v(), tps.L_s(T), tps.p_sat_ice(T), T);
}

// CO.logistic_function                                            CO:124-138
template <class FT> inline FT logistic_function(FT x, FT x_0, FT k) {
    const FT e = eps_numerics<FT>();
    x = jmax(FT(0), x);
    FT x_safe = jmax(x, e);
    FT x_0_safe = jmax(x_0, e);
    FT z = k * (x_safe / x_0_safe - x_0_safe / x_safe);
    FT result = exp_(-log1pexp(-z));
    return (x < e) ? FT(0) : ((x_0 < e) ? FT(1) : result);
}
// CO.logistic_function_integral                                   CO:157-173
template <class FT> inline FT logistic_function_integral(FT x, FT x_0, FT k) {
    const FT e = eps_numerics<FT>();
    x = jmax(FT(0), x);
    FT x_safe = jmax(x, e);
    FT x_0_safe = jmax(x_0, e);
    FT trnslt = -log1mexp(-k) / k;
    FT kt = k * (x_safe / x_0_safe - 1 + trnslt);
    FT result = (log1pexp(kt) / k - trnslt) * x_0_safe;
    return (x < e) ? FT(0) : ((x_0 < e) ? x : result);
}

// UT.fac                                                         UT:304-308
inline double fac(int n) {
    double r = 1;
    for (int i = 2; i <= n; ++i) r *= i;
    return r;
}

// CO.Chen2022_exponential_pdf                                     CO:414-422
template <class FT> inline FT chen2022_exponential_pdf(FT a, FT b, FT c, FT lam_inv, int k) {
    FT delta = FT(k + 1);
    FT gamma_delta = FT(fac(k));
    return a * exp_(-delta * log_(lam_inv) - (b + delta) * log_(1 / lam_inv + c)) *
           tgamma_(b + delta) / gamma_delta;
}

// CO.Chen2022_vel_coeffs(::Chen2022VelTypeRain, ρₐ)                CO:290-300
template <class FT>
inline void chen2022_vel_coeffs_rain(const typename PT<FT>::vel_chen_rain& v, FT rho,
                                     FT aiu[3], FT bi[3], FT ciu[3]) {
    rho = jmax(rho, FT(0));
    FT q = exp_(v.rho0 * rho);
    FT ai[3] = {v.a[0] * q, v.a[1] * q, v.a[2] * q * pow_(rho, v.a3_pow)};
    for (int i = 0; i < 3; ++i) {
        bi[i] = v.b[i] - v.b_rho * rho;
        aiu[i] = ai[i] * pow_(FT(1000), bi[i]);
        ciu[i] = v.c[i] * 1000;
    }
}

// DT.generalized_gamma_Mⁿ                                          DT:109-112
template <class FT> inline FT generalized_gamma_Mn(FT nu, FT mu, FT B, FT N, FT n) {
    return N * pow_(B, -n / mu) * tgamma_((nu + 1 + n) / mu) / tgamma_((nu + 1) / mu);
}

}  // namespace orc
