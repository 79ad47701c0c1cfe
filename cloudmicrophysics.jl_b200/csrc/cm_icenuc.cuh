// cm_icenuc.cuh — ice-nucleation rates, water activities and ARG2000 aerosol activation.
//
// Device form of src/IceNucleation.jl (IN:44-205, 219-253, 557-584), the water activities of
// src/Common.jl (CO:188-271) and src/AerosolActivation.jl (AA:35-433).  Per-mode quantities
// that depend on parameters only (f_i, g_i, the T-independent part of the critical
// supersaturation, log σ_i factors) are folded into ArgK on the host; per point the ARG2000
// kernel needs 2 exponentials for the two saturation pressures, 1 logarithm + 2 exps per
// mode for the (ζ/η)^p1 and (S_m²/(η+3ζ))^p2 powers, one erf per mode, and the handful of
// sqrt / cbrt of the Korolev-Mazin terms.
#pragma once
#include "cm_thermo.cuh"

namespace cm {

constexpr int kMaxModes = 8;

template <class FT> struct ArgK {
    FT A_coef;             // 2 σ M_w / (ρ_w R)            -> A = A_coef / T        AA:35-40
    FT sm_coef[kMaxModes]; // 2/sqrt(hygro) (A_coef/3/r_dry)^(3/2) -> S_m = sm_coef T^(-3/2)   AA:107-118
    FT f[kMaxModes], g[kMaxModes];                  // f1 exp(f2 ln²σ), g1 + g2 ln σ     AA:175-176
    FT inv_eta_coef[kMaxModes];                     // 2π ρ_w N_i
    FT eta_coef[kMaxModes];                         // 1 / (2π ρ_w N_i)                  (Inf for an empty mode, like the reference's division)
    FT t1_coef[kMaxModes];                          // (2π ρ_w N_i)^p1: (ζ/η_i)^p1 = (ζ γ / sq³)^p1 * t1_coef_i — ONE exponential per point for all modes
    FT inv_sm2_coef[kMaxModes];                     // 1 / sm_coef²: 1 / S_m,i² = inv_sm2_coef_i T³
    FT log_sm_coef[kMaxModes], log_inv_eta_coef[kMaxModes];   // their logarithms (host): one log per mode per point instead of three
    FT log_T_triple;
    FT u_coef[kMaxModes];                           // 2 / (3 √2 ln σ_i)                AA:256
    FT m_fac[kMaxModes];                            // 3 ln σ_i √2 / 2                   AA:320
    FT inv_K_safe, inv_D_safe, ln10, c43pi_rho_w, c43pi_rho_i, four_pi;
    FT Rv_over_Rd;
    FT rho_w_rcp;          // correctly rounded 1 / ρ_w: G / ρ_w as an IEEE quotient through divr_ (cm_math.cuh)
};

template <class FT> __host__ inline ArgK<FT> make_arg_k(const cumicro_params_icenuc_f64& p, bool method_is_f32) {
    ArgK<FT> k{};
    const FT pi = FT(3.141592653589793238462643383279502884L);
    const FT epsn = method_is_f32 ? FT(2.2737367544323206e-13) : FT(2.8126442852362996e-103);
    k.A_coef = FT(2) * p.arg.sigma * p.arg.M_w / p.arg.rho_w / p.arg.R;
    for (int i = 0; i < p.n_modes && i < kMaxModes; ++i) {
        const auto& m = p.modes[i];
        const FT ls = std::log(m.stdev);
        k.sm_coef[i] = FT(2) / std::sqrt(m.hygro) * std::pow(k.A_coef / 3 / m.r_dry, FT(1.5));
        k.f[i] = p.arg.f1 * std::exp(p.arg.f2 * ls * ls);
        k.g[i] = p.arg.g1 + p.arg.g2 * ls;
        k.inv_eta_coef[i] = 2 * pi * p.arg.rho_w * m.N;
        k.log_sm_coef[i] = std::log(k.sm_coef[i]);
        k.log_inv_eta_coef[i] = std::log(k.inv_eta_coef[i]);   // -Inf for an empty mode: η = Inf, both powers vanish
        k.eta_coef[i] = FT(1) / k.inv_eta_coef[i];
        k.t1_coef[i] = std::exp(FT(p.arg.p1) * k.log_inv_eta_coef[i]);
        k.inv_sm2_coef[i] = FT(1) / (k.sm_coef[i] * k.sm_coef[i]);
        k.u_coef[i] = FT(2) / (FT(3) * std::sqrt(FT(2)) * ls);
        k.m_fac[i] = FT(3) * ls * std::sqrt(FT(2)) / 2;
    }
    k.inv_K_safe = FT(1) / std::max(FT(p.aps.K_therm), epsn);
    k.inv_D_safe = FT(1) / std::max(FT(p.aps.D_vapor), epsn);
    k.ln10 = FT(2.302585092994045684017991454684364208L);
    k.c43pi_rho_w = FT(4.0 / 3) * pi * p.arg.rho_w;
    k.c43pi_rho_i = FT(4.0 / 3) * pi * p.arg.rho_i;
    k.four_pi = 4 * pi;
    k.Rv_over_Rd = p.tps.R_v / p.tps.R_d;
    k.rho_w_rcp = rcp_cr_((double)p.arg.rho_w);
    k.log_T_triple = std::log(FT(p.tps.T_triple));
    return k;
}

// 10^x                                                               (Julia `10^x`)
template <class FT> CM_DEV FT exp10_(FT x, FT ln10) { return exp_full_(x * ln10); }

// IN.deposition_J / ABIFM_J / HomIceNucleation.homogeneous_J_{cubic,linear}      IN:92-134, 557-584
template <class FT> CM_DEV FT deposition_J(const cumicro_dust_f64& d, FT da_w, FT ln10) {
    return d.has_deposition ? exp10_(fma_(d.deposition_m, da_w, d.deposition_c) + FT(4), ln10) : FT(0);
}
template <class FT> CM_DEV FT ABIFM_J(const cumicro_dust_f64& d, FT da_w, FT ln10) {
    return d.has_ABIFM ? exp10_(fma_(d.ABIFM_m, da_w, d.ABIFM_c) + FT(4), ln10) : FT(0);
}
template <class FT> CM_DEV FT homogeneous_J_cubic(const cumicro_koop2000_f64& ip, FT da_w, FT ln10, bool& domain_error) {
    domain_error = !(ip.da_w_min <= da_w && da_w <= ip.da_w_max);
    const FT d2 = da_w * da_w;
    const FT logJ = ip.c1 + ip.c2 * da_w - ip.c3 * d2 + ip.c4 * (d2 * da_w);
    return exp10_(logJ + FT(6), ln10);
}
template <class FT> CM_DEV FT homogeneous_J_linear(const cumicro_koop2000_f64& ip, FT da_w, FT ln10) {
    return exp10_(fma_(ip.linear_c2, da_w, ip.linear_c1) + FT(6), ln10);
}

struct ArgOut {
    double S_max;
    double N_act[kMaxModes];
    double M_act[kMaxModes];
    double da_w;     // a_w_eT(p_v, T) - a_w_ice(T)
};

// AA.max_supersaturation + N_activated_per_mode + M_activated_per_mode          AA:138-324
// FAST_ERF: erf_fast_ (cm_math.cuh) instead of the CUDA libm's erf for the activated fractions.
// NM >= 0: the number of modes is known at compile time (= p.n_modes, checked by the caller): the mode loops unroll to exactly NM
// bodies with no trip tests — an unrolled run-time loop keeps all kMaxModes guarded copies in the instruction stream.
template <bool WANT_M, bool FAST_ERF = false, int NM = -1>
CM_DEV ArgOut arg2000(const cumicro_params_icenuc_f64& p, const ThermoK<double>& tk, const ArgK<double>& k, double T, double pr,
                      double w, double q_tot, double q_liq, double q_ice, double N_liq, double N_ice,
                      const ThermoShared<double>* shared = nullptr) {
    using FT = double;
    ArgOut o;
    const auto& ap = p.arg;
    const ThermoShared<FT> th = shared ? *shared : thermo_shared(tk, T);
    const TempState<FT>& ts = th.ts;
    const FT p_vs = th.p_vs_l, p_vs_i = th.p_vs_i, inv_pvs = th.inv_pvs_l, inv_pvs_i = th.inv_pvs_i;
    const FT R_v = tk.R_v;
    const FT R_m = p.tps.R_d * (FT(1) + (k.Rv_over_Rd - FT(1)) * q_tot - k.Rv_over_Rd * (q_liq + q_ice));   // TDI.Rₘ
    const FT cpm = cp_m(tk, q_tot, q_liq, q_ice);
    const FT Lv = latent_heat_vapor(tk, T);
    const FT Ls = latent_heat_sublim(tk, T);
    const FT inv_Rm = rcp_(R_m);
    const FT rho_air = pr * (inv_Rm * ts.inv_T);                             // TDI.air_density = p / (R_m T)
    const FT p_v = (q_tot - q_liq - q_ice) * rho_air * R_v * T;
    const FT pv_over_pvs = p_v * inv_pvs;
    o.da_w = pv_over_pvs - p_vs_i * inv_pvs;                                  // CO.a_w_eT - CO.a_w_ice
    const FT G = divr_(G_func(tk, k.inv_K_safe, k.inv_D_safe, Lv, inv_pvs, ts), ap.rho_w, k.rho_w_rcp);
    const FT inv_cpm = rcp_(cpm), inv_p = rcp_(pr);
    const FT alpha = pv_over_pvs * (Lv * ap.g * tk.inv_R_v * inv_cpm * ts.inv_T * ts.inv_T - ap.g * inv_Rm * ts.inv_T);
    const FT common_g = pv_over_pvs * R_m * Lv * tk.inv_R_v * inv_cpm * ts.inv_T * inv_p;
    const FT gamma = fma_(common_g, Lv, R_v * T * inv_pvs);
    const FT A = k.A_coef * ts.inv_T;
    const FT aw_G = alpha * w * rcp_(G);
    const FT sq = sqrtg_(aw_G);
    const FT zeta = FT(2.0 / 3.0) * A * sq;
    const FT sq3 = sq * sq * sq;
    const FT inv_gamma = rcp_(gamma);
    const FT l_zeta = logp_(zeta);
    // logarithms: log S_m,i = log(sm_coef_i) - 3/2 log T and log η_i = log(sq³/γ) - log(2π ρ_w N_i) come from host-side
    // logarithms of the parameters plus two per-point ones; per mode only log(η_i + 3ζ) remains
    auto log_g = [](FT v) { return (v > FT(2.3e-308) && v < FT(1.7e308)) ? logp_(v) : log_full_(v); };
    const FT l_T32 = FT(-1.5) * (ts.log_Tr + k.log_T_triple);
    const FT eta_common = sq3 * inv_gamma;
    const FT l_eta_common = log_g(eta_common);
    FT l_Sm[kMaxModes];
    FT tmp = FT(0);
    // (ζ/η_i)^p1 = E0 (2π ρ_w N_i)^p1 with E0 = (ζ γ / sq³)^p1: one exponential for all modes (an empty mode has t1_coef = 0)
    const FT E0 = exp_full_(ap.p1 * (l_zeta - l_eta_common));
    const FT T3 = T * T * T;
#pragma unroll
    for (int i = 0; i < (NM >= 0 ? NM : kMaxModes); ++i) {
        if (NM < 0 && i >= p.n_modes) break;
        l_Sm[i] = k.log_sm_coef[i] + l_T32;
        const FT eta = eta_common * k.eta_coef[i];
        // (ζ/η)^p1 and (S_m²/(η+3ζ))^p2
        const FT t1 = (k.t1_coef[i] == FT(0)) ? FT(0) : E0 * k.t1_coef[i];
        const FT t2 = exp_full_(ap.p2 * (FT(2) * l_Sm[i] - log_g(fma_(FT(3), zeta, eta))));
        tmp += (k.inv_sm2_coef[i] * T3) * fma_(k.f[i], t1, k.g[i] * t2);
    }
    const FT S_max_ARG = rsqrtg_(tmp);
    const FT r_liq = (N_liq < tk.eps) ? FT(0) : cbrtg_(rho_air * q_liq * rcp_(N_liq * k.c43pi_rho_w));
    const FT K_liq = k.four_pi * ap.rho_w * N_liq * r_liq * G * gamma;
    const FT gamma_i = fma_(common_g, Ls, R_v * T * inv_pvs);
    const FT r_ice = (N_ice < tk.eps) ? FT(0) : cbrtg_(rho_air * q_ice * rcp_(N_ice * k.c43pi_rho_i));
    const FT rhoGi = G_func(tk, k.inv_K_safe, k.inv_D_safe, Ls, inv_pvs_i, ts);
    const FT xi = p_vs * inv_pvs_i;
    const FT K_ice = k.four_pi * N_ice * r_ice * rhoGi * gamma_i;
    const FT aw = alpha * w;
    const FT S_max = S_max_ARG * (aw - K_ice * (xi - FT(1))) * rcp_(fma_(K_liq + K_ice * xi, S_max_ARG, aw));
    o.S_max = clamp0_(S_max);
    const FT l_smax = log_g(o.S_max);   // -Inf when S_max = 0 (libm path): erf(+Inf) = 1 -> N_act = 0   (AA:256)
#pragma unroll
    for (int i = 0; i < (NM >= 0 ? NM : kMaxModes); ++i) {
        if (NM < 0 && i >= p.n_modes) break;
        const FT lr = l_Sm[i] - l_smax;   // log(S_m / S_max)
        o.N_act[i] = p.modes[i].N * FT(0.5) * (FT(1) - (FAST_ERF ? erf_fast_(k.u_coef[i] * lr) : erf_(k.u_coef[i] * lr)));
        if (WANT_M) o.M_act[i] = p.modes[i].molar_mass_mix * FT(0.5) * erfc_(lr / k.m_fac[i] - k.m_fac[i]);
    }
    return o;
}

}  // namespace cm
