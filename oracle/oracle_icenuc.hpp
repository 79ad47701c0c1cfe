// oracle_icenuc.hpp — CPU restatement of src/IceNucleation.jl (heterogeneous + homogeneous
// nucleation rates, INP concentrations), the water activities of src/Common.jl (CO:188-271)
// and the ARG2000 aerosol activation of src/AerosolActivation.jl (AA:35-433).
// TEST INFRASTRUCTURE ONLY (see oracle_base.hpp).  Operation order follows the Julia source.
#pragma once
#include "oracle_1m.hpp"

namespace orc {

// 10^x as Julia evaluates `10^x` for a float exponent: exp10
template <class FT> inline FT exp10_(FT x) { return pow_(FT(10), x); }

// IN.deposition_J / IN.ABIFM_J                                       IN:92-134
template <class FT, class D> inline FT deposition_J(const D& dust, FT da_w) {
    if (!dust.has_deposition) return FT(0);
    FT logJ = FT(dust.deposition_m) * da_w + FT(dust.deposition_c);
    return exp10_(logJ + 4);
}
template <class FT, class D> inline FT ABIFM_J(const D& dust, FT da_w) {
    if (!dust.has_ABIFM) return FT(0);
    FT logJ = FT(dust.ABIFM_m) * da_w + FT(dust.ABIFM_c);
    return exp10_(logJ + 4);
}
// HomIceNucleation.homogeneous_J_cubic (DomainError outside the range -> status)   IN:557-565
template <class FT, class K> inline FT homogeneous_J_cubic(const K& ip, FT da_w, bool& domain_error) {
    domain_error = !(FT(ip.da_w_min) <= da_w && da_w <= FT(ip.da_w_max));
    FT logJ = FT(ip.c1) + FT(ip.c2) * da_w - FT(ip.c3) * (da_w * da_w) + FT(ip.c4) * (da_w * da_w * da_w);
    return exp10_(logJ + 6);
}
// HomIceNucleation.homogeneous_J_linear                                IN:581-584
template <class FT, class K> inline FT homogeneous_J_linear(const K& ip, FT da_w) {
    FT logJ = FT(ip.linear_c2) * da_w + FT(ip.linear_c1);
    return exp10_(logJ + 6);
}
// IN.dust_activated_number_fraction / MohlerDepositionRate (AssertionError -> status)  IN:44-77
template <class FT, class D, class M> inline FT dust_activated_number_fraction(const D& dust, const M& ip, FT Si, FT T, bool& err) {
    err = !(Si < FT(ip.Si_max));
    FT S0 = (T > FT(ip.T_thr)) ? FT(dust.S0_warm) : FT(dust.S0_cold);
    FT a = (T > FT(ip.T_thr)) ? FT(dust.a_warm) : FT(dust.a_cold);
    return jmax(FT(0), exp_(a * (Si - S0)) - 1);
}
template <class FT, class D, class M>
inline FT MohlerDepositionRate(const D& dust, const M& ip, FT Si, FT T, FT dSi_dt, FT N_aer, bool& err) {
    err = !(Si < FT(ip.Si_max));
    FT a = (T > FT(ip.T_thr)) ? FT(dust.a_warm) : FT(dust.a_cold);
    return jmax(FT(0), N_aer * a * dSi_dt);
}
// IN.P3_deposition_N_i / P3_het_N_i                                    IN:162-205
template <class FT, class M> inline FT P3_deposition_N_i(const M& ip, FT T) {
    FT Tp = jmax(FT(ip.T_dep_thres), T);
    FT Ni = 1000 * FT(ip.c1) * exp_(FT(ip.c2) * (FT(ip.T0) - Tp));
    return (T < FT(ip.T0)) ? Ni : FT(0);
}
template <class FT, class M> inline FT P3_het_N_i(const M& ip, FT T, FT N_l, FT V_l, FT dt) {
    FT Ts = FT(ip.T0) - T;
    return N_l * (1 - exp_(-FT(ip.het_B) * V_l * dt * exp_(FT(ip.het_a) * Ts)));
}
// IN.INP_concentration_frequency                                       IN:219-224
template <class FT, class F> inline FT INP_concentration_frequency(const F& p, FT INPC, FT T) {
    if (T >= FT(p.T_freeze)) return FT(0);
    FT mu = INP_concentration_mean<FT>(p, T);
    FT s2 = FT(p.sigma) * FT(p.sigma);
    FT d = log_(INPC) - mu;
    return exp_(-(d * d) / (2 * s2)) / sqrt_(pi<FT>() * 2 * s2);
}

// CO.H2SO4_soln_saturation_vapor_pressure / a_w_xT / a_w_eT / a_w_ice    CO:188-271
template <class FT, class H> inline FT H2SO4_soln_saturation_vapor_pressure(const H& h, FT x, FT T) {
    FT w_h = FT(h.w_2) * x;
    return exp_(FT(h.c[0]) - FT(h.c[1]) * x + FT(h.c[2]) * x * w_h - FT(h.c[3]) * x * (w_h * w_h) +
                (FT(h.c[4]) + FT(h.c[5]) * x - FT(h.c[6]) * x * w_h) / T) * 100;
}
template <class FT, class H> inline FT a_w_xT(const H& h, const Thermo<FT>& tps, FT x, FT T) {
    return H2SO4_soln_saturation_vapor_pressure<FT>(h, x, T) / tps.p_sat_liq(T);
}
template <class FT> inline FT a_w_eT(const Thermo<FT>& tps, FT e, FT T) { return e / tps.p_sat_liq(T); }
template <class FT> inline FT a_w_ice(const Thermo<FT>& tps, FT T) { return tps.p_sat_ice(T) / tps.p_sat_liq(T); }

// ---- ARG2000 (src/AerosolActivation.jl) ----------------------------------------------------------
// AA.coeff_of_curvature                                                AA:35-40
template <class FT, class A> inline FT coeff_of_curvature(const A& ap, FT T) {
    return FT(2) * FT(ap.sigma) * FT(ap.M_w) / FT(ap.rho_w) / FT(ap.R) / T;
}
// AA.critical_supersaturation (one mode)                               AA:107-118
template <class FT, class A, class Mo> inline FT critical_supersaturation(const A& ap, const Mo& mode, FT T) {
    FT Acurv = coeff_of_curvature<FT>(ap, T);
    return 2 / sqrt_(FT(mode.hygro)) * pow_(Acurv / 3 / FT(mode.r_dry), FT(3.0 / 2));
}
// AA.max_supersaturation                                               AA:138-200
template <class FT, class P>
inline FT max_supersaturation(const P& p, FT T, FT pr, FT w, FT q_tot, FT q_liq, FT q_ice, FT N_liq, FT N_ice) {
    const auto& ap = p.arg;
    Thermo<FT> tps(p.tps);
    FT R_v = tps.R_v();
    FT R_m = tps.R_m(q_tot, q_liq, q_ice);
    FT cp_m = tps.cp_m(q_tot, q_liq, q_ice);
    FT Lv = tps.L_v(T);
    FT rho_air = pr / (R_m * T);                                        // TDI.air_density
    FT p_v = (q_tot - q_liq - q_ice) * rho_air * R_v * T;
    FT p_vs = tps.p_sat_liq(T);
    FT G = G_func_liquid<FT>(p.aps, tps, T) / FT(ap.rho_w);
    FT alpha = p_v / p_vs * (Lv * FT(ap.g) / R_v / cp_m / (T * T) - FT(ap.g) / R_m / T);
    FT gamma = R_v * T / p_vs + p_v / p_vs * R_m * (Lv * Lv) / R_v / cp_m / T / pr;
    FT Acurv = coeff_of_curvature<FT>(ap, T);
    FT zeta = 2 * Acurv / 3 * sqrt_(alpha * w / G);
    FT tmp = FT(0);
    for (int i = 0; i < p.n_modes; ++i) {
        const auto& m = p.modes[i];
        FT Sm = critical_supersaturation<FT>(ap, m, T);
        FT ls = log_(FT(m.stdev));
        FT f = FT(ap.f1) * exp_(FT(ap.f2) * (ls * ls));
        FT g = FT(ap.g1) + FT(ap.g2) * ls;
        FT sq = sqrt_(alpha * w / G);
        FT eta = (sq * sq * sq) / (FT(2 * 3.141592653589793238462643383279502884L) * FT(ap.rho_w) * gamma * FT(m.N));
        tmp += 1 / (Sm * Sm) * (f * pow_(zeta / eta, FT(ap.p1)) + g * pow_((Sm * Sm) / (eta + 3 * zeta), FT(ap.p2)));
    }
    FT S_max_ARG = FT(1) / sqrt_(tmp);
    FT c43pi = FT(4.0 / 3 * 3.141592653589793238462643383279502884L);
    FT r_liq = (N_liq < eps<FT>()) ? FT(0) : FT(cbrt_(rho_air * q_liq / N_liq / FT(ap.rho_w) / c43pi));
    FT K_liq = FT(4 * 3.141592653589793238462643383279502884L) * FT(ap.rho_w) * N_liq * r_liq * G * gamma;
    FT Ls = tps.L_s(T);
    FT gamma_i = R_v * T / p_vs + p_v / p_vs * R_m * Lv * Ls / R_v / cp_m / T / pr;
    FT r_ice = (N_ice < eps<FT>()) ? FT(0) : FT(cbrt_(rho_air * q_ice / N_ice / FT(ap.rho_i) / c43pi));
    FT rhoGi = G_func_ice<FT>(p.aps, tps, T);
    FT xi = tps.p_sat_liq(T) / tps.p_sat_ice(T);
    FT K_ice = FT(4 * 3.141592653589793238462643383279502884L) * N_ice * r_ice * rhoGi * gamma_i;
    FT S_max = S_max_ARG * (alpha * w - K_ice * (xi - FT(1))) / (alpha * w + (K_liq + K_ice * xi) * S_max_ARG);
    return jmax(FT(0), S_max);
}
// AA.N_activated_per_mode / M_activated_per_mode                          AA:235-259, 294-324
template <class FT, class P>
inline void activated_per_mode(const P& p, FT T, FT pr, FT w, FT q_tot, FT q_liq, FT q_ice, FT N_liq, FT N_ice, FT& S_max, FT* N_act,
                               FT* M_act) {
    S_max = max_supersaturation<FT>(p, T, pr, w, q_tot, q_liq, q_ice, N_liq, N_ice);
    for (int i = 0; i < p.n_modes; ++i) {
        const auto& m = p.modes[i];
        FT Sm = critical_supersaturation<FT>(p.arg, m, T);
        FT ls = log_(FT(m.stdev));
        if (N_act) {
            FT u = 2 * log_(Sm / S_max) / 3 / sqrt_(FT(2)) / ls;
            N_act[i] = FT(m.N) * FT(0.5) * (1 - erf_(u));
        }
        if (M_act) {
            FT fac = 3 * ls * sqrt_(FT(2)) / 2;
            FT u = log_(Sm / S_max) / fac;
            M_act[i] = FT(m.molar_mass_mix) / 2 * erfc_(u - fac);
        }
    }
}

}  // namespace orc
