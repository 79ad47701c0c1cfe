// oracle_1m.hpp — CPU restatement of the 1-moment scheme: src/Microphysics1M.jl,
// the non-equilibrium cloud formation of src/MicrophysicsNonEq.jl, and the 1-moment
// part of src/BulkMicrophysicsTendencies.jl (BMT:141-632).  TEST INFRASTRUCTURE ONLY
// (see oracle_base.hpp).  Operation order follows the Julia source line by line.
#pragma once
#include "oracle_2m.hpp"

namespace orc {

template <class FT> struct PT1 {
    using params = typename std::conditional<std::is_same<FT, float>::value, cumicro_params_1m_f32, cumicro_params_1m_f64>::type;
    using mass = typename std::conditional<std::is_same<FT, float>::value, cumicro_particle_mass_f32, cumicro_particle_mass_f64>::type;
    using frost = typename std::conditional<std::is_same<FT, float>::value, cumicro_frostenberg2023_f32, cumicro_frostenberg2023_f64>::type;
};

// ---- CM1.get_n0 / get_v0 / lambda_inverse                         CM1:83-152
template <class FT> inline FT get_n0_snow(FT mu, FT nu, FT q_sno, FT rho) {
    const FT e = eps_numerics<FT>();
    FT safe_q = jmax(q_sno, e);
    return (q_sno > e) ? FT(mu * pow_(rho * safe_q, FT(nu))) : FT(0);
}
template <class FT, class VR> inline FT get_v0_rain(const VR& vel, FT rho) {
    FT density_factor = jmax(FT(vel.rho_w) / rho - 1, FT(0));
    return sqrt_(FT(8.0 / 3) / FT(vel.C_drag) * density_factor * FT(vel.grav) * FT(vel.r0));
}
template <class FT, class M> inline FT lambda_inverse(FT n0, const M& mass, FT q, FT rho) {
    const FT e = eps_numerics<FT>();
    FT qp = clamp_to_nonneg(q);
    FT rhop = clamp_to_nonneg(rho);
    FT denom = FT(mass.chi_m) * FT(mass.m0) * jmax(n0, e) * FT(mass.gamma_coeff);
    FT lam_inv = pow_(rhop * qp * pow_(FT(mass.r0), FT(mass.me + mass.dm)) / denom, FT(1 / (mass.me + mass.dm + 1)));
    return jmax(FT(mass.r0) * FT(1e-5), lam_inv);
}

// CM1.size_distr_parameters                                          CM1:375-388
template <class FT> struct SizeDistr1M { FT lam_rai, n0_rai, v0_rai, lam_sno, n0_sno, v0_sno, lam_icl, n0_icl; };
template <class FT, class P>
inline SizeDistr1M<FT> size_distr_parameters(const P& mp, FT q_rai, FT q_sno, FT q_icl, FT rho) {
    SizeDistr1M<FT> sd;
    sd.n0_rai = FT(mp.rain.n0);
    sd.lam_rai = lambda_inverse<FT>(sd.n0_rai, mp.rain.mass, q_rai, rho);
    sd.v0_rai = get_v0_rain<FT>(mp.vel_rain, rho);
    sd.n0_sno = get_n0_snow<FT>(FT(mp.snow.mu), FT(mp.snow.nu), q_sno, rho);
    sd.lam_sno = lambda_inverse<FT>(sd.n0_sno, mp.snow.mass, q_sno, rho);
    sd.v0_sno = FT(mp.vel_snow.v0);
    sd.n0_icl = FT(mp.cloud_ice.n0);
    sd.lam_icl = lambda_inverse<FT>(sd.n0_icl, mp.cloud_ice.mass, q_icl, rho);
    return sd;
}

// CM1.terminal_velocity (Blk1M, with v0 and λ⁻¹ given)                CM1:223-238
template <class FT, class V, class M>
inline FT terminal_velocity_blk1m(const V& vel, const M& mass, FT q, FT v0, FT lam_inv) {
    FT fall_w = FT(vel.chi_v) * v0 * pow_(lam_inv / FT(mass.r0), FT(vel.ve + vel.dv)) * FT(vel.gamma_term) / FT(mass.gamma_coeff);
    return (q > eps_numerics<FT>()) ? fall_w : FT(0);
}

// IN.INP_concentration_mean                                          IN:250-253
template <class FT, class F> inline FT INP_concentration_mean(const F& ip, FT T) {
    FT T_celsius = jmin(T - FT(ip.T_freeze), FT(0));
    return 9 * log_(-FT(ip.b) * T_celsius / 10) - FT(ip.log_a);
}
// NEQ.τ_relax                                                        NEQ:32-50
template <class FT, class F> inline FT tau_relax_frostenberg(FT rho_i, FT D_vapor, const F& ip, FT q_icl, FT T) {
    const FT e = eps_numerics<FT>();
    FT N_icl = exp_(INP_concentration_mean<FT>(ip, T));
    FT safe_N = jmax(N_icl, e);
    FT r = (N_icl > e) ? FT(cbrt_((3 * q_icl) / (4 * pi<FT>() * safe_N * rho_i))) : FT(0);
    FT r0 = FT(1e-6);
    FT r_safe = jmax(r, r0);
    return 1 / (4 * pi<FT>() * D_vapor * N_icl * r_safe);
}

// NEQ.conv_q_vap_to_q_icl(::TemperatureDependent)                     NEQ:195-224
template <class FT, class P>
inline FT conv_q_vap_to_q_icl_tempdep(const P& mp, const Thermo<FT>& tps, FT q_tot, FT q_lcl, FT q_icl, FT q_rai, FT q_sno,
                                      FT rho, FT T) {
    FT tau_sub = FT(mp.pp.cloud_ice_tau_relax);
    FT tau_dep = tau_relax_frostenberg<FT>(FT(mp.cloud_ice.rho_i), FT(mp.aps.D_vapor), mp.pp.frostenberg, q_icl, T);
    FT Rv = tps.R_v();
    FT Ls = tps.L_s(T);
    FT cp_air = tps.cp_m(q_tot, q_lcl + q_rai, q_icl + q_sno);
    FT qv = Thermo<FT>::q_vap(q_tot, q_lcl + q_rai, q_icl + q_sno);
    FT qv_sat_ice = tps.q_sat_ice(T, rho);
    FT dqsi_dT = dqcld_dT(qv_sat_ice, Ls, Rv, T);
    FT Gam = gamma_helper(Ls, cp_air, dqsi_dT);
    FT sat_excess = qv - qv_sat_ice;
    FT tendency = (sat_excess < 0) ? FT(-jmin(-sat_excess, jmax(FT(0), q_icl)) / (tau_sub * Gam)) : FT(sat_excess / (tau_dep * Gam));
    bool limiter = (T > tps.T_freeze()) && (tendency > FT(0));
    return limiter ? FT(0) : tendency;
}

// CM1.warm_accretion_melt_factor                                       CM1:458-465
template <class FT> inline FT warm_accretion_melt_factor(const Thermo<FT>& tps, FT T) {
    FT L_f = tps.L_f(T);
    FT dT = T - tps.T_freeze();
    return (T <= tps.T_freeze()) ? FT(0) : FT(tps.cv_l() / L_f * dT);
}

// CM1.accretion (low-level kernel)                                     CM1:491-514
template <class FT, class V, class M, class A>
inline FT accretion_1m(const V& vel, const M& mass, const A& area, FT E, FT q_clo, FT q_pre, FT n0, FT v0, FT lam_inv) {
    const FT e = eps_numerics<FT>();
    FT rate = q_clo * E * n0 * FT(area.a0) * v0 * FT(area.chi_a) * FT(vel.chi_v) * lam_inv * FT(vel.gamma_accr) /
              pow_(FT(mass.r0) / lam_inv, FT(area.ae + vel.ve + area.da + vel.dv));
    return (q_clo > e && q_pre > e) ? rate : FT(0);
}

// CM1.accretion_rain_sink                                              CM1:535-561
template <class FT, class P>
inline FT accretion_rain_sink(const P& mp, FT E, FT q_icl, FT q_rai, FT rho, FT n0_ice, FT lam_ice_inv, FT n0, FT v0, FT lam_inv) {
    const FT e = eps_numerics<FT>();
    const auto& m = mp.rain.mass;
    const auto& a = mp.rain.area;
    const auto& v = mp.vel_rain;
    FT rate = E / rho * n0 * n0_ice * FT(m.m0) * FT(a.a0) * v0 * FT(m.chi_m) * FT(a.chi_a) * FT(v.chi_v) * lam_ice_inv * lam_inv *
              FT(v.gamma_accr_rain_sink) / pow_(FT(m.r0) / lam_inv, FT(m.me + a.ae + v.ve + m.dm + a.da + v.dv));
    return (q_icl > e && q_rai > e) ? rate : FT(0);
}

// CM1.accretion_snow_rain (low-level kernel)                           CM1:604-644
template <class FT, class Mj>
inline FT accretion_snow_rain_kernel(const Mj& mass_j, FT v_ti, FT v_tj, FT E_ij, FT coeff_disp, FT q_i, FT q_j, FT rho,
                                     FT n0_i, FT n0_j, FT lam_i, FT lam_j) {
    const FT e = eps_numerics<FT>();
    FT delta = FT(mass_j.me + mass_j.dm);
    FT dv = v_ti - v_tj;
    FT dv_eff = sqrt_(dv * dv + coeff_disp * (v_ti * v_ti + v_tj * v_tj));
    FT rate = pi<FT>() / rho * n0_i * n0_j * FT(mass_j.m0) * FT(mass_j.chi_m) * E_ij * dv_eff * FT(mass_j.gamma_coeff) /
              pow_(FT(mass_j.r0), delta) *
              (2 * (lam_i * lam_i * lam_i) * pow_(lam_j, delta + 1) + 2 * (delta + 1) * (lam_i * lam_i) * pow_(lam_j, delta + 2) +
               (delta + 2) * (delta + 1) * lam_i * pow_(lam_j, delta + 3));
    return (q_i > e && q_j > e) ? rate : FT(0);
}

// ventilated diffusional growth factor shared by CM1:915-960, 990-1037, 1092-1139:
//   a_vent + b_vent cbrt(Sc) / (r0/λ⁻¹)^((ve+Δv)/2) sqrt(2 v0 χv/ν_air λ⁻¹) gamma_vent
template <class FT, class V, class M, class Ve, class Air>
inline FT vent_factor_1m(const V& vel, const M& mass, const Ve& vent, const Air& aps, FT v0, FT lam_inv) {
    FT Sc = FT(aps.nu_air) / jmax(FT(aps.D_vapor), eps_numerics<FT>());
    return FT(vent.a) + FT(vent.b) * cbrt_(Sc) / pow_(FT(mass.r0) / lam_inv, FT((vel.ve + vel.dv) / 2)) *
                            sqrt_(2 * v0 * FT(vel.chi_v) / FT(aps.nu_air) * lam_inv) * FT(vel.gamma_vent);
}

// ---- BMT._microphysics_source_terms                                  BMT:141-217
enum {
    S1M_PHASE_VAP_LCL = 0, S1M_PHASE_VAP_ICL, S1M_ACNV_LCL_RAI, S1M_ACNV_ICL_SNO, S1M_ACCR_LCL_RAI, S1M_ACCR_LCL_SNO_COLD,
    S1M_ACCR_LCL_SNO_WARM, S1M_ACCR_MELT_LCL_SNO, S1M_ACCR_ICL_RAI, S1M_ACCR_FREEZE_ICL_RAI, S1M_ACCR_ICL_SNO,
    S1M_ACCR_RAI_SNO_COLD, S1M_ACCR_RAI_SNO_WARM, S1M_ACCR_MELT_RAI_SNO, S1M_PHASE_VAP_RAI, S1M_PHASE_VAP_SNO, S1M_MELT_ICL_LCL,
    S1M_MELT_SNO_RAI, S1M_NSRC
};

template <class FT> struct Src1M { FT s[S1M_NSRC]; };

template <class FT, class P>
inline Src1M<FT> microphysics_source_terms_1m(const P& mp, FT rho, FT T, FT q_tot, FT q_lcl, FT q_icl, FT q_rai, FT q_sno) {
    rho = clamp_to_nonneg(rho);
    q_tot = clamp_to_nonneg(q_tot);
    q_lcl = clamp_to_nonneg(q_lcl);
    q_icl = clamp_to_nonneg(q_icl);
    q_rai = clamp_to_nonneg(q_rai);
    q_sno = clamp_to_nonneg(q_sno);
    const FT e = eps_numerics<FT>();
    Thermo<FT> tps(mp.tps);
    const auto& o = mp.processes;
    const auto& pp = mp.pp;
    Src1M<FT> r;
    SizeDistr1M<FT> sd = size_distr_parameters<FT>(mp, q_rai, q_sno, q_icl, rho);
    const FT T_freeze = tps.T_freeze();

    // phase change vapour <-> cloud                                    BMT:165-167
    r.s[S1M_PHASE_VAP_LCL] = o.cloud_liquid_formation
                                 ? conv_q_vap_to_q_lcl_const<FT>(FT(pp.cloud_liquid_tau_relax), tps, q_tot, q_lcl, q_icl, q_rai, q_sno, rho, T)
                                 : FT(0);
    r.s[S1M_PHASE_VAP_ICL] = (o.cloud_ice_formation == CUMICRO_1M_CLOUD_ICE_CONSTANT_TIMESCALE)
                                 ? conv_q_vap_to_q_icl_const<FT>(FT(pp.cloud_ice_tau_relax), tps, q_tot, q_lcl, q_icl, q_rai, q_sno, rho, T)
                                 : ((o.cloud_ice_formation == CUMICRO_1M_CLOUD_ICE_TEMPERATURE_DEPENDENT)
                                        ? conv_q_vap_to_q_icl_tempdep<FT>(mp, tps, q_tot, q_lcl, q_icl, q_rai, q_sno, rho, T)
                                        : FT(0));
    // autoconversion                                                    CM1:352-364, 412-446
    if (o.rain_autoconversion == CUMICRO_1M_RAIN_ACNV_KESSLER)
        r.s[S1M_ACNV_LCL_RAI] = logistic_function_integral<FT>(q_lcl, FT(pp.rain_acnv_q_threshold), FT(pp.rain_acnv_k)) / FT(pp.rain_acnv_tau);
    else if (o.rain_autoconversion == CUMICRO_1M_RAIN_ACNV_PRESCRIBED_ND)
        r.s[S1M_ACNV_LCL_RAI] = jmax(FT(0), q_lcl) / (FT(pp.rain_acnv_tau) * pow_(FT(pp.rain_acnv_Nc) / FT(100000000), FT(pp.rain_acnv_alpha)));
    else
        r.s[S1M_ACNV_LCL_RAI] = FT(0);
    if (o.snow_autoconversion == CUMICRO_1M_SNOW_ACNV_NO_SUPERSAT)
        r.s[S1M_ACNV_ICL_SNO] = logistic_function_integral<FT>(q_icl, FT(pp.snow_acnv_q_threshold), FT(pp.snow_acnv_k)) / FT(pp.snow_acnv_tau);
    else if (o.snow_autoconversion == CUMICRO_1M_SNOW_ACNV_WITH_SUPERSAT) {
        FT r_is = FT(pp.snow_acnv_r_ice_snow);
        FT S = tps.supersat_ice(q_tot, q_lcl + q_rai, q_icl + q_sno, rho, T);
        FT G = G_func_ice<FT>(mp.aps, tps, T);
        FT lam = sd.lam_icl;
        FT med = FT(mp.cloud_ice.mass.me + mp.cloud_ice.mass.dm);
        FT rate = 4 * pi<FT>() * S * G * sd.n0_icl / rho * exp_(-r_is / lam) * (r_is * r_is / med + (r_is / lam + 1) * (lam * lam));
        r.s[S1M_ACNV_ICL_SNO] = (q_icl > e && S > FT(0) && T < T_freeze) ? rate : FT(0);
    } else
        r.s[S1M_ACNV_ICL_SNO] = FT(0);

    const bool is_warm = T >= T_freeze;                                // BMT:174
    // accretions                                                        BMT:176-204
    r.s[S1M_ACCR_LCL_RAI] = o.cloud_liquid_rain_accretion
                                ? accretion_1m<FT>(mp.vel_rain, mp.rain.mass, mp.rain.area, FT(pp.e_lcl_rai), q_lcl, q_rai, sd.n0_rai, sd.v0_rai, sd.lam_rai)
                                : FT(0);
    {
        FT S = FT(0), S_melt = FT(0);
        if (o.cloud_liquid_snow_accretion) {
            S = accretion_1m<FT>(mp.vel_snow, mp.snow.mass, mp.snow.area, FT(pp.e_lcl_sno), q_lcl, q_sno, sd.n0_sno, sd.v0_sno, sd.lam_sno);
            S_melt = warm_accretion_melt_factor<FT>(tps, T) * S;
        }
        r.s[S1M_ACCR_LCL_SNO_COLD] = is_warm ? FT(0) : S;
        r.s[S1M_ACCR_LCL_SNO_WARM] = is_warm ? S : FT(0);
        r.s[S1M_ACCR_MELT_LCL_SNO] = S_melt;
    }
    r.s[S1M_ACCR_ICL_RAI] = o.cloud_ice_rain_accretion
                                ? accretion_1m<FT>(mp.vel_rain, mp.rain.mass, mp.rain.area, FT(pp.e_icl_rai), q_icl, q_rai, sd.n0_rai, sd.v0_rai, sd.lam_rai)
                                : FT(0);
    r.s[S1M_ACCR_FREEZE_ICL_RAI] = o.cloud_ice_rain_accretion
                                       ? accretion_rain_sink<FT>(mp, FT(pp.e_icl_rai), q_icl, q_rai, rho, sd.n0_icl, sd.lam_icl, sd.n0_rai, sd.v0_rai, sd.lam_rai)
                                       : FT(0);
    r.s[S1M_ACCR_ICL_SNO] = o.cloud_ice_snow_accretion
                                ? accretion_1m<FT>(mp.vel_snow, mp.snow.mass, mp.snow.area, FT(pp.e_icl_sno), q_icl, q_sno, sd.n0_sno, sd.v0_sno, sd.lam_sno)
                                : FT(0);
    {
        FT S_rai_sno = FT(0), S_sno_rai = FT(0), S_melt = FT(0);
        if (o.rain_snow_accretion) {
            FT v_sno = terminal_velocity_blk1m<FT>(mp.vel_snow, mp.snow.mass, q_sno, sd.v0_sno, sd.lam_sno);
            FT v_rai = terminal_velocity_blk1m<FT>(mp.vel_rain, mp.rain.mass, q_rai, sd.v0_rai, sd.lam_rai);
            // (type_i, type_j) = (snow, rain): rain freezes on snow            CM1:831-848
            S_rai_sno = accretion_snow_rain_kernel<FT>(mp.rain.mass, v_sno, v_rai, FT(pp.e_rai_sno), FT(pp.coeff_disp), q_sno, q_rai, rho,
                                                       sd.n0_sno, sd.n0_rai, sd.lam_sno, sd.lam_rai);
            // (type_i, type_j) = (rain, snow)                                   CM1:849-865
            S_sno_rai = accretion_snow_rain_kernel<FT>(mp.snow.mass, v_rai, v_sno, FT(pp.e_rai_sno), FT(pp.coeff_disp), q_rai, q_sno, rho,
                                                       sd.n0_rai, sd.n0_sno, sd.lam_rai, sd.lam_sno);
            S_melt = warm_accretion_melt_factor<FT>(tps, T) * S_rai_sno;
        }
        r.s[S1M_ACCR_RAI_SNO_COLD] = is_warm ? FT(0) : S_rai_sno;
        r.s[S1M_ACCR_RAI_SNO_WARM] = is_warm ? S_sno_rai : FT(0);
        r.s[S1M_ACCR_MELT_RAI_SNO] = is_warm ? S_melt : FT(0);
    }
    // precipitation <-> vapour                                          CM1:915-1037
    if (o.rain_condensation_evaporation) {
        FT S = tps.supersat_liq(q_tot, q_lcl + q_rai, q_icl + q_sno, rho, T);
        FT G = G_func_liquid<FT>(mp.aps, tps, T);
        FT lam = sd.lam_rai;
        FT rate = 4 * pi<FT>() * sd.n0_rai / rho * S * G * (lam * lam) *
                  vent_factor_1m<FT>(mp.vel_rain, mp.rain.mass, mp.rain.vent, mp.aps, sd.v0_rai, lam);
        bool cond = q_rai > e && S < FT(0);
        r.s[S1M_PHASE_VAP_RAI] = jmin(FT(0), cond ? rate : FT(0));
    } else
        r.s[S1M_PHASE_VAP_RAI] = FT(0);
    if (o.snow_deposition_sublimation) {
        FT S = tps.supersat_ice(q_tot, q_lcl + q_rai, q_icl + q_sno, rho, T);
        FT G = G_func_ice<FT>(mp.aps, tps, T);
        FT lam = sd.lam_sno;
        FT rate = 4 * pi<FT>() * sd.n0_sno / rho * S * G * (lam * lam) *
                  vent_factor_1m<FT>(mp.vel_snow, mp.snow.mass, mp.snow.vent, mp.aps, sd.v0_sno, lam);
        FT v = (q_sno > e) ? rate : FT(0);
        r.s[S1M_PHASE_VAP_SNO] = (o.snow_deposition_sublimation == CUMICRO_1M_SNOW_SUBLIMATION_ONLY) ? jmin(FT(0), v) : v;
    } else
        r.s[S1M_PHASE_VAP_SNO] = FT(0);
    // melting                                                            CM1:1053-1139
    if (o.cloud_ice_melt) {
        FT L = tps.L_f(T);
        FT lam = sd.lam_icl;
        FT rate = 4 * pi<FT>() * FT(mp.cloud_ice.n0) / rho * FT(mp.aps.K_therm) / L * (T - T_freeze) * (lam * lam);
        r.s[S1M_MELT_ICL_LCL] = (q_icl > e && T > T_freeze) ? rate : FT(0);
    } else
        r.s[S1M_MELT_ICL_LCL] = FT(0);
    if (o.snow_melt) {
        FT L = tps.L_f(T);
        FT lam = sd.lam_sno;
        FT rate = 4 * pi<FT>() * sd.n0_sno / rho * FT(mp.aps.K_therm) / L * (T - T_freeze) * (lam * lam) *
                  vent_factor_1m<FT>(mp.vel_snow, mp.snow.mass, mp.snow.vent, mp.aps, sd.v0_sno, lam);
        r.s[S1M_MELT_SNO_RAI] = (q_sno > e && T > T_freeze) ? rate : FT(0);
    } else
        r.s[S1M_MELT_SNO_RAI] = FT(0);
    return r;
}

// BMT._aggregate_tendencies                                            BMT:227-252
template <class FT> inline void aggregate_tendencies_1m(const Src1M<FT>& r, FT out[4]) {
    const FT* s = r.s;
    out[0] = s[S1M_PHASE_VAP_LCL] - s[S1M_ACNV_LCL_RAI] - s[S1M_ACCR_LCL_RAI] - s[S1M_ACCR_LCL_SNO_COLD] - s[S1M_ACCR_LCL_SNO_WARM] +
             s[S1M_MELT_ICL_LCL];
    out[1] = s[S1M_PHASE_VAP_ICL] - s[S1M_ACNV_ICL_SNO] - s[S1M_ACCR_ICL_RAI] - s[S1M_ACCR_ICL_SNO] - s[S1M_MELT_ICL_LCL];
    out[2] = s[S1M_ACNV_LCL_RAI] + s[S1M_ACCR_LCL_RAI] + s[S1M_ACCR_LCL_SNO_WARM] + s[S1M_ACCR_MELT_LCL_SNO] - s[S1M_ACCR_FREEZE_ICL_RAI] -
             s[S1M_ACCR_RAI_SNO_COLD] + s[S1M_ACCR_RAI_SNO_WARM] + s[S1M_ACCR_MELT_RAI_SNO] + s[S1M_PHASE_VAP_RAI] + s[S1M_MELT_SNO_RAI];
    out[3] = s[S1M_ACNV_ICL_SNO] + s[S1M_ACCR_LCL_SNO_COLD] - s[S1M_ACCR_MELT_LCL_SNO] + s[S1M_ACCR_ICL_RAI] + s[S1M_ACCR_FREEZE_ICL_RAI] +
             s[S1M_ACCR_ICL_SNO] + s[S1M_ACCR_RAI_SNO_COLD] - s[S1M_ACCR_RAI_SNO_WARM] - s[S1M_ACCR_MELT_RAI_SNO] + s[S1M_PHASE_VAP_SNO] -
             s[S1M_MELT_SNO_RAI];
}

// BMT._linearize + _linearized_implicit_step                            BMT:269-465
template <class FT> inline FT muladd_(FT a, FT b, FT c);
template <> inline double muladd_(double a, double b, double c) { return std::fma(a, b, c); }
template <> inline float muladd_(float a, float b, float c) { return std::fma(a, b, c); }
template <> inline Tr muladd_(Tr a, Tr b, Tr c) { return a * b + c; }

template <class FT, class P>
inline void linearized_implicit_step_1m(const P& mp, FT rho, FT T, FT q_tot, FT q_lcl, FT q_icl, FT q_rai, FT q_sno, FT dt, FT out[4]) {
    Src1M<FT> r = microphysics_source_terms_1m<FT>(mp, rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno);
    const FT* s = r.s;
    Thermo<FT> tps(mp.tps);
    const FT q_min = FT(mp.tps.q_min);
    FT M11 = 0, M12 = 0, M22 = 0, M31 = 0, M33 = 0, M34 = 0, M41 = 0, M42 = 0, M43 = 0, M44 = 0, e1 = 0, e2 = 0, e4 = 0;
    FT D;
    bool is_source;
    D = s[S1M_PHASE_VAP_LCL] / jmax(q_min, q_lcl);
    is_source = s[S1M_PHASE_VAP_LCL] >= FT(0);
    e1 += is_source ? s[S1M_PHASE_VAP_LCL] : FT(0);
    M11 += is_source ? FT(0) : D;
    D = s[S1M_PHASE_VAP_ICL] / jmax(q_min, q_icl);
    is_source = s[S1M_PHASE_VAP_ICL] >= FT(0);
    e2 += is_source ? s[S1M_PHASE_VAP_ICL] : FT(0);
    M22 += is_source ? FT(0) : D;
    D = s[S1M_MELT_ICL_LCL] / jmax(q_min, q_icl);
    M22 -= D;
    M12 += D;
    D = s[S1M_ACNV_LCL_RAI] / jmax(q_min, q_lcl);
    M11 -= D;
    M31 += D;
    D = s[S1M_ACNV_ICL_SNO] / jmax(q_min, q_icl);
    M22 -= D;
    M42 += D;
    D = s[S1M_ACCR_LCL_RAI] / jmax(q_min, q_lcl);
    M11 -= D;
    M31 += D;
    FT D_cold = s[S1M_ACCR_LCL_SNO_COLD] / jmax(q_min, q_lcl);
    FT D_warm = s[S1M_ACCR_LCL_SNO_WARM] / jmax(q_min, q_lcl);
    M11 -= D_cold + D_warm;
    M31 += D_warm;
    M41 += D_cold;
    D = s[S1M_ACCR_MELT_LCL_SNO] / jmax(q_min, q_sno);
    M44 -= D;
    M34 += D;
    D = s[S1M_ACCR_ICL_RAI] / jmax(q_min, q_icl);
    M22 -= D;
    M42 += D;
    D = s[S1M_ACCR_ICL_SNO] / jmax(q_min, q_icl);
    M22 -= D;
    M42 += D;
    D = s[S1M_ACCR_FREEZE_ICL_RAI] / jmax(q_min, q_rai);
    M33 -= D;
    M43 += D;
    D = s[S1M_ACCR_RAI_SNO_WARM] / jmax(q_min, q_sno);
    M44 -= D;
    M34 += D;
    D = s[S1M_ACCR_MELT_RAI_SNO] / jmax(q_min, q_sno);
    M44 -= D;
    M34 += D;
    D = s[S1M_ACCR_RAI_SNO_COLD] / jmax(q_min, q_rai);
    M33 -= D;
    M43 += D;
    D = (-s[S1M_PHASE_VAP_RAI]) / jmax(q_min, q_rai);
    M33 -= D;
    D = s[S1M_PHASE_VAP_SNO] / jmax(q_min, q_sno);
    is_source = s[S1M_PHASE_VAP_SNO] >= FT(0);
    e4 += is_source ? s[S1M_PHASE_VAP_SNO] : FT(0);
    M44 += is_source ? FT(0) : D;
    D = s[S1M_MELT_SNO_RAI] / jmax(q_min, q_sno);
    M44 -= D;
    M34 += D;

    FT inv_dt = FT(1) / dt;
    FT q_sat_min = jmin(tps.q_sat_liq(T, rho), tps.q_sat_ice(T, rho));
    FT q_v = q_tot - q_lcl - q_icl - q_rai - q_sno;
    FT alpha = jmin(FT(1), jmax(FT(0), q_v - q_sat_min) * inv_dt / jmax(e1 + e2 + e4, eps<FT>()));
    FT a11 = inv_dt - M11, a12 = -M12, a22 = inv_dt - M22, a31 = -M31, a33 = inv_dt - M33, a34 = -M34, a41 = -M41, a42 = -M42,
       a43 = -M43, a44 = inv_dt - M44;
    FT b1 = alpha * e1 + inv_dt * q_lcl;
    FT b2 = alpha * e2 + inv_dt * q_icl;
    FT b3 = inv_dt * q_rai;
    FT b4 = alpha * e4 + inv_dt * q_sno;
    FT det12 = a11 * a22;
    FT q_lcl_new = (b1 * a22 - a12 * b2) / det12;
    FT q_icl_new = a11 * b2 / det12;
    FT r3 = muladd_<FT>(-a31, q_lcl_new, b3);
    FT r4 = muladd_<FT>(-a41, q_lcl_new, muladd_<FT>(-a42, q_icl_new, b4));
    FT det = muladd_<FT>(-a34, a43, a33 * a44);
    FT q_rai_new = (r3 * a44 - a34 * r4) / det;
    FT q_sno_new = (a33 * r4 - r3 * a43) / det;
    out[0] = (q_lcl_new - q_lcl) * inv_dt;
    out[1] = (q_icl_new - q_icl) * inv_dt;
    out[2] = (q_rai_new - q_rai) * inv_dt;
    out[3] = (q_sno_new - q_sno) * inv_dt;
}

// BMT.bulk_microphysics_tendencies(::LinearizedAverage, ...)             BMT:572-632
template <class FT, class P>
inline void bmt1m_linearized_average(const P& mp, FT rho, FT T, FT q_tot, FT q_lcl, FT q_icl, FT q_rai, FT q_sno, FT dt, int nsub,
                                     FT out[4]) {
    FT q0[4] = {q_lcl, q_icl, q_rai, q_sno};
    FT dt_sub = dt / FT(nsub);
    FT Lv_over_cp = FT(mp.tps.LH_v0) / FT(mp.tps.cp_d);
    FT Ls_over_cp = FT(mp.tps.LH_s0) / FT(mp.tps.cp_d);
    for (int it = 0; it < nsub; ++it) {
        FT rt[4];
        linearized_implicit_step_1m<FT>(mp, rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno, dt_sub, rt);
        q_lcl += rt[0] * dt_sub;
        q_icl += rt[1] * dt_sub;
        q_rai += rt[2] * dt_sub;
        q_sno += rt[3] * dt_sub;
        T += (Lv_over_cp * (rt[0] + rt[2]) + Ls_over_cp * (rt[1] + rt[3])) * dt_sub;
    }
    out[0] = (q_lcl - q0[0]) / dt;
    out[1] = (q_icl - q0[1]) / dt;
    out[2] = (q_rai - q0[2]) / dt;
    out[3] = (q_sno - q0[3]) / dt;
}

// ---- terminal velocities ------------------------------------------------------------------
// CO.Chen2022_vel_coeffs(::Chen2022VelTypeSmallIce, ρₐ, ρᵢ)              CO:302-324
template <class FT, class V> inline void chen2022_vel_coeffs_small_ice(const V& c, FT rho_a, FT rho_i, FT aiu[2], FT bi[2], FT ciu[2]) {
    rho_a = jmax(rho_a, FT(0));
    FT l = log_(rho_i), sq = sqrt_(rho_i);
    FT As = FT(c.A[1]) * (l * l) - FT(c.A[2]) * l + FT(c.A[0]);
    FT Bs = 1 / (FT(c.B[0]) + FT(c.B[1]) * l + FT(c.B[2]) / sq);
    FT Cs = FT(c.C[0]) + FT(c.C[1]) * exp_(FT(c.C[2]) * rho_i) + FT(c.C[3]) * sq;
    FT Es = FT(c.E[0]) - FT(c.E[1]) * (l * l) + FT(c.E[2]) * sq;
    FT Fs = -exp_(FT(c.F[0]) - FT(c.F[1]) * (l * l) + FT(c.F[2]) * l);
    FT Gs = 1 / (FT(c.G[0]) + FT(c.G[1]) / l - FT(c.G[2]) * l / rho_i);
    FT ai[2] = {Es * pow_(rho_a, As), Fs * pow_(rho_a, As)};
    bi[0] = bi[1] = Bs + rho_a * Cs;
    FT ci[2] = {FT(0), Gs};
    for (int i = 0; i < 2; ++i) {
        aiu[i] = ai[i] * pow_(FT(1000), bi[i]);
        ciu[i] = ci[i] * 1000;
    }
}
// CO.Chen2022_vel_coeffs(::Chen2022VelTypeLargeIce, ρₐ, ρᵢ)              CO:326-349
template <class FT, class V> inline void chen2022_vel_coeffs_large_ice(const V& c, FT rho_a, FT rho_i, FT aiu[2], FT bi[2], FT ciu[2]) {
    rho_a = jmax(rho_a, FT(0));
    FT l = log_(rho_i), sq = sqrt_(rho_i);
    FT Al = FT(c.A[0]) + FT(c.A[1]) * l + FT(c.A[2]) / (rho_i * sq);
    FT Bl = exp_(FT(c.B[0]) + FT(c.B[1]) * (l * l) + FT(c.B[2]) * l);
    FT Cl = exp_(FT(c.C[0]) + FT(c.C[1]) / l + FT(c.C[2]) / rho_i);
    FT El = FT(c.E[0]) + FT(c.E[1]) * l * sq + FT(c.E[2]) * sq;
    FT Fl = FT(c.F[0]) + FT(c.F[1]) * l - exp_(log_(FT(-c.F[2])) - rho_i);
    FT Gl = 1 / (FT(c.G[0]) + FT(c.G[1]) * l * sq + FT(c.G[2]) / sq);
    FT Hl = FT(c.H[0]) + FT(c.H[1]) * (rho_i * rho_i) * sq + exp_(log_(FT(-c.H[2])) - rho_i);
    FT ai[2] = {Bl * pow_(rho_a, Al), El * pow_(rho_a, Al) * exp_(Hl * rho_a)};
    bi[0] = Cl;
    bi[1] = Fl;
    FT ci[2] = {FT(0), Gl};
    for (int i = 0; i < 2; ++i) {
        aiu[i] = ai[i] * pow_(FT(1000), bi[i]);
        ciu[i] = ci[i] * 1000;
    }
}

// CM1.terminal_velocity(::Rain|::Snow, ::Blk1MVelType, ρ, q)              CM1:240-249
template <class FT, class P> inline FT terminal_velocity_1m_rain_blk(const P& mp, FT rho, FT q) {
    FT v0 = get_v0_rain<FT>(mp.vel_rain, rho);
    FT lam = lambda_inverse<FT>(FT(mp.rain.n0), mp.rain.mass, q, rho);
    return terminal_velocity_blk1m<FT>(mp.vel_rain, mp.rain.mass, q, v0, lam);
}
template <class FT, class P> inline FT terminal_velocity_1m_snow_blk(const P& mp, FT rho, FT q) {
    FT n0 = get_n0_snow<FT>(FT(mp.snow.mu), FT(mp.snow.nu), q, rho);
    FT lam = lambda_inverse<FT>(n0, mp.snow.mass, q, rho);
    return terminal_velocity_blk1m<FT>(mp.vel_snow, mp.snow.mass, q, FT(mp.vel_snow.v0), lam);
}
// CM1.terminal_velocity(::Rain, ::Chen2022VelTypeRain, ρ, q)               CM1:251-270
template <class FT, class P, class V> inline FT terminal_velocity_1m_rain_chen(const P& mp, const V& vel, FT rho, FT q) {
    FT aiu[3], bi[3], ciu[3];
    chen2022_vel_coeffs_rain<FT>(vel, rho, aiu, bi, ciu);
    FT lam_r = lambda_inverse<FT>(FT(mp.rain.n0), mp.rain.mass, q, rho);
    FT lam_d = 2 * lam_r;
    FT w = chen2022_exponential_pdf<FT>(aiu[0], bi[0], ciu[0], lam_d, 3) + chen2022_exponential_pdf<FT>(aiu[1], bi[1], ciu[1], lam_d, 3) +
           chen2022_exponential_pdf<FT>(aiu[2], bi[2], ciu[2], lam_d, 3);
    w = jmax(FT(0), w);
    return (q > eps_numerics<FT>()) ? w : FT(0);
}
// CM1.terminal_velocity(::Snow, ::Chen2022VelTypeLargeIce, ρ, q)           CM1:272-291
template <class FT, class P, class V> inline FT terminal_velocity_1m_snow_chen(const P& mp, const V& vel, FT rho, FT q) {
    FT aiu[2], bi[2], ciu[2];
    chen2022_vel_coeffs_large_ice<FT>(vel, rho, FT(mp.snow.rho_i), aiu, bi, ciu);
    FT n0 = get_n0_snow<FT>(FT(mp.snow.mu), FT(mp.snow.nu), q, rho);
    FT lam_d = 2 * lambda_inverse<FT>(n0, mp.snow.mass, q, rho);
    FT pk = pow_(FT(mp.snow.aspr_phi), FT(mp.snow.aspr_kappa));
    FT w = pk * chen2022_exponential_pdf<FT>(aiu[0], bi[0], ciu[0], lam_d, 3) + pk * chen2022_exponential_pdf<FT>(aiu[1], bi[1], ciu[1], lam_d, 3);
    w = jmax(FT(0), w);
    return (q > eps_numerics<FT>()) ? w : FT(0);
}
// NEQ.terminal_velocity(::CloudLiquid, ::StokesRegimeVelType, ρₐ, q)        NEQ:250-262
template <class FT, class P, class V> inline FT terminal_velocity_noneq_liquid(const P& mp, const V& vel, FT rho, FT q) {
    FT pref = FT(1.0 / 18) * (FT(vel.rho_w) / rho - 1) * FT(vel.grav) / FT(vel.nu_air);
    FT safe_q = clamp_to_nonneg(q);
    FT D = cbrt_(FT(6 / 3.141592653589793238462643383279502884L) * rho * safe_q / FT(mp.cloud_liquid.N_0) / FT(mp.cloud_liquid.rho_w));
    FT w = pref * (D * D);
    return (q > eps_numerics<FT>()) ? w : FT(0);
}
// NEQ.terminal_velocity(::CloudIce, ::Chen2022VelTypeSmallIce, ρₐ, q)        NEQ:264-281
template <class FT, class P, class V> inline FT terminal_velocity_noneq_ice(const P& mp, const V& vel, FT rho, FT q) {
    FT aiu[2], bi[2], ciu[2];
    chen2022_vel_coeffs_small_ice<FT>(vel, rho, FT(mp.cloud_ice.rho_i), aiu, bi, ciu);
    FT safe_q = clamp_to_nonneg(q);
    FT D = cbrt_(FT(6 / 3.141592653589793238462643383279502884L) * rho * safe_q / FT(mp.cloud_ice.N_0) / FT(mp.cloud_ice.rho_i));
    FT v = FT(0);
    for (int i = 0; i < 2; ++i) v += aiu[i] * pow_(D, bi[i]) * exp_(-ciu[i] * D);   // Chen2022VelocityCurve  CO:391-392
    FT w = jmax(FT(0), v);
    return (q > eps_numerics<FT>()) ? w : FT(0);
}

}  // namespace orc
