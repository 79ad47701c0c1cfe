// cm_1m.cuh — 1-moment bulk microphysics, fused per grid point.
//
// Device form of BMT._microphysics_source_terms / _aggregate_tendencies /
// _linearized_implicit_step (BMT:141-465) and their callees in src/Microphysics1M.jl
// (CM1:83-152, 223-249, 352-1139) and src/MicrophysicsNonEq.jl (NEQ:32-224).
//
// The reference evaluates ~20 real powers per point (lambda_inverse, snow n0, the
// (r0/λ⁻¹)^x factors of every accretion / ventilation / fall-speed formula).  All of them
// are powers of three per-point quantities — ρ q_rai, ρ q_sno, ρ q_icl — so this kernel takes
// ONE logarithm per species and forms every power as exp_(x · log λ⁻¹ + const) with the
// parameter-only parts folded into host-side constants (OneMK).  The saturation vapour
// pressures, latent heats and Marshall-Palmer parameters are computed once and shared by
// all 18 source terms.  Regime predicates compare the same quantities with the same
// operators as the reference.
#pragma once
#include "cm_thermo.cuh"

namespace cm {

enum {
    S1M_PHASE_VAP_LCL = 0, S1M_PHASE_VAP_ICL, S1M_ACNV_LCL_RAI, S1M_ACNV_ICL_SNO, S1M_ACCR_LCL_RAI, S1M_ACCR_LCL_SNO_COLD,
    S1M_ACCR_LCL_SNO_WARM, S1M_ACCR_MELT_LCL_SNO, S1M_ACCR_ICL_RAI, S1M_ACCR_FREEZE_ICL_RAI, S1M_ACCR_ICL_SNO,
    S1M_ACCR_RAI_SNO_COLD, S1M_ACCR_RAI_SNO_WARM, S1M_ACCR_MELT_RAI_SNO, S1M_PHASE_VAP_RAI, S1M_PHASE_VAP_SNO, S1M_MELT_ICL_LCL,
    S1M_MELT_SNO_RAI, S1M_NSRC
};

// Host-derived constants of one Marshall-Palmer species.
template <class FT> struct MPSpeciesK {
    FT log_coef;     // log(r0^(me+Δm) / (χm m0 Γ(me+Δm+1)))   [n0 enters per point for snow]
    FT inv_exp;      // 1/(me+Δm+1)
    FT lam_floor;    // r0 1e-5
    FT log_lam_floor;
    FT log_r0;       // log(mass.r0)
};

template <class FT> struct OneMK {
    MPSpeciesK<FT> rai, sno, icl;
    FT log_n0_rai, log_n0_icl, log_mu_sno, log_eps_numerics;
    FT v0_rai_pref;          // 8/3/C_drag grav vel_rain.r0
    // accretion onto rain / snow: n0 a0 χa χv Γaccr (n0 of snow per point), exponent x = ae+ve+Δa+Δv
    FT accr_rai_pref, accr_rai_x, accr_sno_pref, accr_sno_x;
    FT sink_pref, sink_x;    // n0_rai n0_icl m0 a0 χm χa χv Γsink; me+ae+ve+Δm+Δa+Δv
    // Blk1M fall speeds: χv Γterm/Γcoeff and exponent ve+Δv
    FT vt_rai_pref, vt_rai_x, vt_sno_pref, vt_sno_x;
    // rain-snow collisions, arm j = rain / snow: π m0_j χm_j E Γcoeff_j / r0_j^δ_j, δ_j
    FT rs_rai_pref, rs_rai_delta, rs_sno_pref, rs_sno_delta, coeff_disp;
    // ventilation: b cbrt(Sc) Γvent sqrt(2 χv/ν_air) and exponent (ve+Δv)/2
    FT vent_rai_a, vent_rai_b, vent_rai_x, vent_sno_a, vent_sno_b, vent_sno_x;
    FT four_pi, inv_K_safe, inv_D_safe;
    // logistic autoconversion: trnslt = -log1mexp(-k)/k
    FT rain_trnslt, snow_trnslt, rain_inv_q_thr, snow_inv_q_thr;
    FT prescribed_nd_inv;    // 1/(τ (Nc/1e8)^α)
    FT ice_med;              // cloud-ice me+Δm (WithSupersaturation)
    FT frost_c, frost_r0;    // 4 π D_vapor; FT(1e-6)
    // IEEE reciprocals of uniform divisors whose numerators are exact zeros at gated-off points (divr_, cm_math.cuh)
    FT rain_inv_k, snow_inv_k, rain_inv_tau, snow_inv_tau;
    // default exponent structure (onem_std_exponents): every power of λ⁻¹ the body needs is an integer power of
    // u_r = (λ_r/r0)^(1/4) resp. u_s = (λ_s/r0)^(1/8) — r0 and the powers of it that go with them
    FT r0_rai, r0_rai4, sqrt_r0_rai, r0_sno, r0_sno3, sqrt_r0_sno;
    FT inv_cloud_ice_tau;    // 1 / τ_relax of cloud ice
    FT rs_r_c1, rs_r_c2, rs_s_c1, rs_s_c2;   // 2 (δ + 1) and (δ + 2)(δ + 1) of the two collision arms (the per-point expressions, hoisted: same bits)
    int std_exponents;
};

// The reference's default process options (Microphysics1MOptions.jl: every process on, snow deposition AND sublimation): with STD
// they are compile-time constants — no option loads, no option branches.
struct DefaultProcesses1M {
    static constexpr int cloud_liquid_formation = 1, cloud_ice_formation = 1, cloud_ice_melt = 1, rain_autoconversion = 1,
                         snow_autoconversion = 1, rain_condensation_evaporation = 1, snow_deposition_sublimation = 2, snow_melt = 1,
                         cloud_liquid_rain_accretion = 1, cloud_liquid_snow_accretion = 1, cloud_ice_rain_accretion = 1,
                         cloud_ice_snow_accretion = 1, rain_snow_accretion = 1;
};
template <class O> __host__ inline bool processes_are_default(const O& o) {
    using D = DefaultProcesses1M;
    return o.cloud_liquid_formation == D::cloud_liquid_formation && o.cloud_ice_formation == D::cloud_ice_formation &&
           o.cloud_ice_melt == D::cloud_ice_melt && o.rain_autoconversion == D::rain_autoconversion &&
           o.snow_autoconversion == D::snow_autoconversion && o.rain_condensation_evaporation == D::rain_condensation_evaporation &&
           o.snow_deposition_sublimation == D::snow_deposition_sublimation && o.snow_melt == D::snow_melt &&
           o.cloud_liquid_rain_accretion == D::cloud_liquid_rain_accretion && o.cloud_liquid_snow_accretion == D::cloud_liquid_snow_accretion &&
           o.cloud_ice_rain_accretion == D::cloud_ice_rain_accretion && o.cloud_ice_snow_accretion == D::cloud_ice_snow_accretion &&
           o.rain_snow_accretion == D::rain_snow_accretion;
}
template <bool STD, class PARAMS> CM_DEV auto processes_of(const PARAMS& p) {
    if constexpr (STD) return DefaultProcesses1M{};
    else return p.processes;
}
template <class FT> __host__ inline OneMK<FT> make_1m_k(const typename P<FT>::params_1m& p, bool method_is_f32 = false) {
    OneMK<FT> k{};
    const FT pi = FT(3.141592653589793238462643383279502884L);
    const FT epsn = method_is_f32 ? FT(2.2737367544323206e-13) : FT(2.8126442852362996e-103);
    // FT(1e-5), FT(1e-6) of the method's float type (CM1:151, NEQ:46)
    const FT c1em5 = method_is_f32 ? FT(1e-5f) : FT(1e-5);
    k.frost_r0 = method_is_f32 ? FT(1e-6f) : FT(1e-6);
    auto species = [&](const typename P<FT>::particle_mass& m) {
        MPSpeciesK<FT> s;
        const FT d = m.me + m.dm;
        s.log_coef = std::log(std::pow(m.r0, d) / (m.chi_m * m.m0 * m.gamma_coeff));
        s.inv_exp = FT(1) / (d + 1);
        s.lam_floor = method_is_f32 ? FT(float(m.r0) * 1e-5f) : m.r0 * c1em5;
        s.log_lam_floor = std::log(s.lam_floor);
        s.log_r0 = std::log(m.r0);
        return s;
    };
    k.rai = species(p.rain.mass);
    k.sno = species(p.snow.mass);
    k.icl = species(p.cloud_ice.mass);
    k.log_n0_rai = std::log(std::max(p.rain.n0, epsn));
    k.log_n0_icl = std::log(std::max(p.cloud_ice.n0, epsn));
    k.log_mu_sno = std::log(p.snow.mu);
    k.log_eps_numerics = std::log(epsn);
    k.v0_rai_pref = FT(8.0 / 3) / p.vel_rain.C_drag * p.vel_rain.grav * p.vel_rain.r0;
    k.accr_rai_pref = p.rain.n0 * p.rain.area.a0 * p.rain.area.chi_a * p.vel_rain.chi_v * p.vel_rain.gamma_accr;
    k.accr_rai_x = p.rain.area.ae + p.vel_rain.ve + p.rain.area.da + p.vel_rain.dv;
    k.accr_sno_pref = p.snow.area.a0 * p.snow.area.chi_a * p.vel_snow.chi_v * p.vel_snow.gamma_accr * p.vel_snow.v0;
    k.accr_sno_x = p.snow.area.ae + p.vel_snow.ve + p.snow.area.da + p.vel_snow.dv;
    k.sink_pref = p.rain.n0 * p.cloud_ice.n0 * p.rain.mass.m0 * p.rain.area.a0 * p.rain.mass.chi_m * p.rain.area.chi_a *
                  p.vel_rain.chi_v * p.vel_rain.gamma_accr_rain_sink * p.pp.e_icl_rai;
    k.sink_x = p.rain.mass.me + p.rain.area.ae + p.vel_rain.ve + p.rain.mass.dm + p.rain.area.da + p.vel_rain.dv;
    k.vt_rai_pref = p.vel_rain.chi_v * p.vel_rain.gamma_term / p.rain.mass.gamma_coeff;
    k.vt_rai_x = p.vel_rain.ve + p.vel_rain.dv;
    k.vt_sno_pref = p.vel_snow.chi_v * p.vel_snow.v0 * p.vel_snow.gamma_term / p.snow.mass.gamma_coeff;
    k.vt_sno_x = p.vel_snow.ve + p.vel_snow.dv;
    k.rs_rai_delta = p.rain.mass.me + p.rain.mass.dm;
    k.rs_sno_delta = p.snow.mass.me + p.snow.mass.dm;
    k.rs_rai_pref = pi * p.rain.mass.m0 * p.rain.mass.chi_m * p.pp.e_rai_sno * p.rain.mass.gamma_coeff / std::pow(p.rain.mass.r0, k.rs_rai_delta);
    k.rs_sno_pref = pi * p.snow.mass.m0 * p.snow.mass.chi_m * p.pp.e_rai_sno * p.snow.mass.gamma_coeff / std::pow(p.snow.mass.r0, k.rs_sno_delta);
    k.coeff_disp = p.pp.coeff_disp;
    const FT cbrt_Sc = std::cbrt(p.aps.nu_air / std::max(p.aps.D_vapor, epsn));
    k.vent_rai_a = p.rain.vent.a;
    k.vent_rai_b = p.rain.vent.b * cbrt_Sc * p.vel_rain.gamma_vent * std::sqrt(2 * p.vel_rain.chi_v / p.aps.nu_air);
    k.vent_rai_x = (p.vel_rain.ve + p.vel_rain.dv) / 2;
    k.vent_sno_a = p.snow.vent.a;
    k.vent_sno_b = p.snow.vent.b * cbrt_Sc * p.vel_snow.gamma_vent * std::sqrt(2 * p.vel_snow.v0 * p.vel_snow.chi_v / p.aps.nu_air);
    k.vent_sno_x = (p.vel_snow.ve + p.vel_snow.dv) / 2;
    k.four_pi = 4 * pi;
    k.inv_K_safe = FT(1) / std::max(p.aps.K_therm, epsn);
    k.inv_D_safe = FT(1) / std::max(p.aps.D_vapor, epsn);
    auto trn = [](FT kk) {  // -log1mexp(-k)/k, LogExpFunctions.log1mexp
        const FT x = -kk;
        const FT l = (x < FT(-0.6931471805599453)) ? std::log1p(-std::exp(x)) : std::log(-std::expm1(x));
        return -l / kk;
    };
    k.rain_trnslt = trn(p.pp.rain_acnv_k);
    k.snow_trnslt = trn(p.pp.snow_acnv_k);
    k.rain_inv_q_thr = FT(1) / std::max(p.pp.rain_acnv_q_threshold, epsn);
    k.snow_inv_q_thr = FT(1) / std::max(p.pp.snow_acnv_q_threshold, epsn);
    k.prescribed_nd_inv = FT(1) / (p.pp.rain_acnv_tau * std::pow(p.pp.rain_acnv_Nc / FT(100000000), p.pp.rain_acnv_alpha));
    k.ice_med = p.cloud_ice.mass.me + p.cloud_ice.mass.dm;
    k.frost_c = 4 * pi * p.aps.D_vapor;
    k.inv_cloud_ice_tau = FT(1) / p.pp.cloud_ice_tau_relax;
    k.rs_r_c1 = FT(2) * (k.rs_rai_delta + FT(1)); k.rs_r_c2 = (k.rs_rai_delta + FT(2)) * (k.rs_rai_delta + FT(1));
    k.rs_s_c1 = FT(2) * (k.rs_sno_delta + FT(1)); k.rs_s_c2 = (k.rs_sno_delta + FT(2)) * (k.rs_sno_delta + FT(1));
    k.r0_rai = p.rain.mass.r0; k.r0_rai4 = (k.r0_rai * k.r0_rai) * (k.r0_rai * k.r0_rai); k.sqrt_r0_rai = std::sqrt(k.r0_rai);
    k.r0_sno = p.snow.mass.r0; k.r0_sno3 = k.r0_sno * k.r0_sno * k.r0_sno; k.sqrt_r0_sno = std::sqrt(k.r0_sno);
    // rain: me+Δm = 3, ae+ve+Δ = 2.5, ve+Δv = 0.5; snow: me+Δm = 2, ae+ve+Δ = 2.25, ve+Δv = 0.25 (the reference's default 1-moment
    // parameters): quarter resp. eighth powers only
    k.std_exponents = (k.accr_rai_x == FT(2.5) && k.sink_x == FT(5.5) && k.vt_rai_x == FT(0.5) && k.rs_rai_delta == FT(3) &&
                       k.vent_rai_x == FT(0.25) && k.accr_sno_x == FT(2.25) && k.vt_sno_x == FT(0.25) && k.rs_sno_delta == FT(2) &&
                       k.vent_sno_x == FT(0.125) && processes_are_default(p.processes)) ? 1 : 0;
    k.rain_inv_k = FT(1) / p.pp.rain_acnv_k; k.snow_inv_k = FT(1) / p.pp.snow_acnv_k;
    k.rain_inv_tau = FT(1) / p.pp.rain_acnv_tau; k.snow_inv_tau = FT(1) / p.pp.snow_acnv_tau;
    return k;
}

// LogExpFunctions.log1pexp (same branch cuts as the package for Float64 / Float32)
CM_DEV double log1pexp_(double x) {
    if (x < -36.7368005696771) return exp_full_(x);
    if (x < 18.021826694558577) return log1p_pos_(exp_(x));   // e^x in [1e-16, 6.7e7]
    if (x < 33.23111882352963) return x + exp_(-x);
    return x;
}
CM_DEV float log1pexp_(float x) {
    if (x < -15.942385f) return expf(x);
    if (x < 9.011913f) return log1pf(expf(x));
    if (x < 16.635532f) return x + expf(-x);
    return x;
}

// CO.logistic_function_integral                                     CO:157-173
template <class FT> CM_DEV FT logistic_function_integral(FT e, FT x, FT x_0, FT inv_x0_safe, FT k, FT inv_k, FT trnslt) {
    x = clamp0_(x);
    const FT x_safe = fmax_(x, e);
    const FT x0_safe = fmax_(x_0, e);
    const FT kt = k * (x_safe * inv_x0_safe - FT(1) + trnslt);
    const FT result = (divr_(log1pexp_(kt), k, inv_k) - trnslt) * x0_safe;
    return (x < e) ? FT(0) : ((x_0 < e) ? x : result);
}

template <class FT> struct Src1M { FT s[S1M_NSRC]; };

// log λ⁻¹ alone (floored like λ⁻¹ itself)
template <class FT> CM_DEV void lambda_inverse_log(const MPSpeciesK<FT>& sk, FT log_rho_q, FT log_n0, FT& loglam) {
    const FT ll = (log_rho_q + sk.log_coef - log_n0) * sk.inv_exp;
    loglam = !(ll > sk.log_lam_floor) ? sk.log_lam_floor : ll;
}
// λ⁻¹ of one species and its logarithm                              CM1:126-152
template <class FT> CM_DEV void lambda_inverse(const MPSpeciesK<FT>& sk, FT log_rho_q, FT log_n0, FT& lam, FT& loglam) {
    const FT ll = (log_rho_q + sk.log_coef - log_n0) * sk.inv_exp;
    const bool floored = !(ll > sk.log_lam_floor);
    loglam = floored ? sk.log_lam_floor : ll;
    lam = floored ? sk.lam_floor : exp_(ll);
}

// STD: the block has the default STRUCTURE (OneMK::std_exponents, decided on the host: default exponents AND default process options): the ~11 real powers of λ_r⁻¹ and
// λ_s⁻¹ (accretion, sink, fall speeds, collision arms, ventilation) are integer powers of ONE exponential per species,
// u_r = (λ_r/r0)^(1/4) and u_s = (λ_s/r0)^(1/8), formed by ~16 multiplications instead of 9 further exp_ calls (~24 instructions
// each).  Same quantities to rounding (a product of <= 26 factors of u: <= 3e-15 relative).
template <class FT, bool STD = false>
CM_DEV Src1M<FT> microphysics_source_terms_1m(const typename P<FT>::params_1m& p, const ThermoK<FT>& tk, const OneMK<FT>& k,
                                              FT rho, FT T, FT q_tot, FT q_lcl, FT q_icl, FT q_rai, FT q_sno,
                                              const ThermoShared<FT>* shared = nullptr) {
    const FT e = tk.eps_n;
    const auto o = processes_of<STD>(p);      // STD: the default process options as compile-time constants
    const auto& pp = p.pp;
    Src1M<FT> r;
    rho = clamp0_(rho);                                           // BMT:146-151
    q_tot = clamp0_(q_tot);
    q_lcl = clamp0_(q_lcl);
    q_icl = clamp0_(q_icl);
    q_rai = clamp0_(q_rai);
    q_sno = clamp0_(q_sno);
    const FT inv_rho = rcp_(rho);
    const FT T_freeze = tk.T_freeze;

    // ---- thermodynamic state (shared)
    const ThermoShared<FT> th = shared ? *shared : thermo_shared(tk, T);   // (e = tk.eps_n)
    const TempState<FT>& ts = th.ts;
    const FT p_vs_l = th.p_vs_l, p_vs_i = th.p_vs_i, inv_pvs_l = th.inv_pvs_l, inv_pvs_i = th.inv_pvs_i;
    const FT Lv = latent_heat_vapor(tk, T);
    const FT Ls = latent_heat_sublim(tk, T);
    const FT Lf = latent_heat_fusion(tk, T);
    const FT q_liq = q_lcl + q_rai;
    const FT q_ice = q_icl + q_sno;
    const FT qv = q_vap(q_tot, q_liq, q_ice);
    const FT rho_Rv_T = rho * tk.R_v * T;
    const FT inv_rho_Rv_T = rcp_(rho_Rv_T);
    const FT qv_sat_l = p_vs_l * inv_rho_Rv_T;
    const FT qv_sat_i = p_vs_i * inv_rho_Rv_T;
    const FT cp_air = cp_m(tk, q_tot, q_liq, q_ice);
    const FT S_liq = fma_(qv * rho_Rv_T, inv_pvs_l, FT(-1));          // TDI.supersaturation_over_liquid
    const FT S_ice = fma_(qv * rho_Rv_T, inv_pvs_i, FT(-1));          // TDI.supersaturation_over_ice
    const FT G_liq = G_func(tk, k.inv_K_safe, k.inv_D_safe, Lv, inv_pvs_l, ts);
    const FT G_ice = G_func(tk, k.inv_K_safe, k.inv_D_safe, Ls, inv_pvs_i, ts);

    // ---- CM1.size_distr_parameters                                   CM1:375-388
    FT lam_r, ll_r, lam_s, ll_s, lam_i, ll_i;
    lambda_inverse(k.icl, logp_(rho * q_icl), k.log_n0_icl, lam_i, ll_i);
    const FT L_sno = logp_(rho * q_sno);
    const bool has_sno = q_sno > e;
    const FT log_n0_s = has_sno ? fma_(p.snow.nu, L_sno, k.log_mu_sno) : k.log_eps_numerics;
    const FT n0_s = has_sno ? exp_(log_n0_s) : FT(0);                  // CM1.get_n0 (snow)
    // powers of u_r = (λ_r/r0)^(1/4), u_s = (λ_s/r0)^(1/8) (STD only)
    FT ur2 = 0, ur3 = 0, ur10 = 0, ur16 = 0, ur22 = 0, us2 = 0, us5 = 0, us18 = 0, us24 = 0;
    if constexpr (STD) {
        lambda_inverse_log(k.rai, logp_(rho * q_rai), k.log_n0_rai, ll_r);
        lambda_inverse_log(k.sno, L_sno, log_n0_s, ll_s);
        const FT ur = exp_(FT(0.25) * (ll_r - k.rai.log_r0));
        ur2 = ur * ur; ur3 = ur2 * ur;
        const FT ur4 = ur2 * ur2, ur8 = ur4 * ur4;
        ur10 = ur8 * ur2; ur16 = ur8 * ur8; ur22 = ur16 * (ur4 * ur2);
        lam_r = k.r0_rai * ur4;
        const FT us = exp_(FT(0.125) * (ll_s - k.sno.log_r0));
        us2 = us * us;
        const FT us4 = us2 * us2, us8 = us4 * us4, us16 = us8 * us8;
        us5 = us4 * us; us18 = us16 * us2; us24 = us16 * us8;
        lam_s = k.r0_sno * us8;
    } else {
        lambda_inverse(k.rai, logp_(rho * q_rai), k.log_n0_rai, lam_r, ll_r);
        lambda_inverse(k.sno, L_sno, log_n0_s, lam_s, ll_s);
    }
    const FT v0_r = sqrtg_(k.v0_rai_pref * fmax_(p.vel_rain.rho_w * inv_rho - FT(1), FT(0)));   // CM1.get_v0 (rain)
    const FT dl_r = ll_r - k.rai.log_r0, dl_s = ll_s - k.sno.log_r0;   // log(λ⁻¹/r0)

    // ---- phase change vapour <-> cloud                               NEQ:110-224
    if (o.cloud_liquid_formation) {
        const FT dqs = qv_sat_l * fma_(Lv * tk.inv_R_v * ts.inv_T, ts.inv_T, -ts.inv_T);
        const FT inv_ts = cp_air * rcp_(pp.cloud_liquid_tau_relax * fma_(Lv, dqs, cp_air));
        const FT se = qv - qv_sat_l;
        r.s[S1M_PHASE_VAP_LCL] = (se < FT(0)) ? -fmin_(-se, q_lcl) * inv_ts : se * inv_ts;
    } else
        r.s[S1M_PHASE_VAP_LCL] = FT(0);
    if (o.cloud_ice_formation) {
        const FT dqs = qv_sat_i * fma_(Ls * tk.inv_R_v * ts.inv_T, ts.inv_T, -ts.inv_T);
        const FT inv_gam = cp_air * rcp_(fma_(Ls, dqs, cp_air));      // 1/Γᵢ
        const FT se = qv - qv_sat_i;
        FT inv_tau_dep = k.inv_cloud_ice_tau;
        const FT inv_tau_sub = inv_tau_dep;
        if (o.cloud_ice_formation == CUMICRO_1M_CLOUD_ICE_TEMPERATURE_DEPENDENT) {
            // NEQ.τ_relax (Frostenberg 2023 INP number, monodisperse radius)   NEQ:32-50, IN:250-253
            const FT Tc = fmin_(T - pp.frostenberg.T_freeze, FT(0));
            const FT N_icl = exp_full_(FT(9) * log_full_(-pp.frostenberg.b * Tc / FT(10)) - pp.frostenberg.log_a);
            const FT safe_N = fmax_(N_icl, e);
            const FT rr = (N_icl > e) ? cbrt_full_((FT(3) * q_icl) / (k.four_pi * safe_N * p.cloud_ice.rho_i)) : FT(0);
            inv_tau_dep = k.frost_c * N_icl * fmax_(rr, k.frost_r0);
        }
        const FT tend = (se < FT(0)) ? -fmin_(-se, q_icl) * inv_tau_sub * inv_gam : se * inv_tau_dep * inv_gam;
        r.s[S1M_PHASE_VAP_ICL] = ((T > T_freeze) && (tend > FT(0))) ? FT(0) : tend;   // NEQ.INP_limiter
    } else
        r.s[S1M_PHASE_VAP_ICL] = FT(0);

    // ---- autoconversion                                               CM1:352-364, 412-446
    if (o.rain_autoconversion == CUMICRO_1M_RAIN_ACNV_KESSLER)
        r.s[S1M_ACNV_LCL_RAI] =
            divr_(logistic_function_integral<FT>(e, q_lcl, pp.rain_acnv_q_threshold, k.rain_inv_q_thr, pp.rain_acnv_k, k.rain_inv_k, k.rain_trnslt),
                  pp.rain_acnv_tau, k.rain_inv_tau);
    else if (o.rain_autoconversion == CUMICRO_1M_RAIN_ACNV_PRESCRIBED_ND)
        r.s[S1M_ACNV_LCL_RAI] = q_lcl * k.prescribed_nd_inv;
    else
        r.s[S1M_ACNV_LCL_RAI] = FT(0);
    if (o.snow_autoconversion == CUMICRO_1M_SNOW_ACNV_NO_SUPERSAT)
        r.s[S1M_ACNV_ICL_SNO] =
            divr_(logistic_function_integral<FT>(e, q_icl, pp.snow_acnv_q_threshold, k.snow_inv_q_thr, pp.snow_acnv_k, k.snow_inv_k, k.snow_trnslt),
                  pp.snow_acnv_tau, k.snow_inv_tau);
    else if (o.snow_autoconversion == CUMICRO_1M_SNOW_ACNV_WITH_SUPERSAT) {
        const FT r_is = pp.snow_acnv_r_ice_snow;
        const FT x = r_is * rcp_(lam_i);
        const FT rate = k.four_pi * S_ice * G_ice * p.cloud_ice.n0 * inv_rho * exp_full_(-x) *
                        fma_(x + FT(1), lam_i * lam_i, r_is * r_is / k.ice_med);
        r.s[S1M_ACNV_ICL_SNO] = (q_icl > e && S_ice > FT(0) && T < T_freeze) ? rate : FT(0);
    } else
        r.s[S1M_ACNV_ICL_SNO] = FT(0);

    const bool is_warm = T >= T_freeze;                                 // BMT:174
    // CM1.warm_accretion_melt_factor                                    CM1:458-465
    const FT alpha_melt = (T <= T_freeze) ? FT(0) : tk.cv_l * rcp_(Lf) * (T - T_freeze);

    // ---- accretion of cloud condensate by rain / snow                 CM1:491-514, 707-810
    {
        // n0 a0 v0 χa χv λ⁻¹ Γaccr / (r0/λ⁻¹)^x  (without q_clo E)
        const FT base_r = k.accr_rai_pref * v0_r * lam_r * (STD ? ur10 : exp_(k.accr_rai_x * dl_r));
        const FT base_s = k.accr_sno_pref * n0_s * lam_s * (STD ? us18 : exp_(k.accr_sno_x * dl_s));
        const bool rai_on = q_rai > e, sno_on = q_sno > e, lcl_on = q_lcl > e, icl_on = q_icl > e;
        r.s[S1M_ACCR_LCL_RAI] = (o.cloud_liquid_rain_accretion && lcl_on && rai_on) ? q_lcl * pp.e_lcl_rai * base_r : FT(0);
        const FT S_ls = (o.cloud_liquid_snow_accretion && lcl_on && sno_on) ? q_lcl * pp.e_lcl_sno * base_s : FT(0);
        r.s[S1M_ACCR_LCL_SNO_COLD] = is_warm ? FT(0) : S_ls;
        r.s[S1M_ACCR_LCL_SNO_WARM] = is_warm ? S_ls : FT(0);
        r.s[S1M_ACCR_MELT_LCL_SNO] = alpha_melt * S_ls;
        r.s[S1M_ACCR_ICL_RAI] = (o.cloud_ice_rain_accretion && icl_on && rai_on) ? q_icl * pp.e_icl_rai * base_r : FT(0);
        r.s[S1M_ACCR_ICL_SNO] = (o.cloud_ice_snow_accretion && icl_on && sno_on) ? q_icl * pp.e_icl_sno * base_s : FT(0);
        // CM1.accretion_rain_sink                                        CM1:535-561
        const FT sink = k.sink_pref * inv_rho * v0_r * lam_i * lam_r * (STD ? ur22 : exp_(k.sink_x * dl_r));
        r.s[S1M_ACCR_FREEZE_ICL_RAI] = (o.cloud_ice_rain_accretion && icl_on && rai_on) ? sink : FT(0);
    }

    // ---- rain-snow collisions                                          CM1:604-644, 812-867
    if (o.rain_snow_accretion) {
        const FT v_r = (q_rai > e) ? k.vt_rai_pref * v0_r * (STD ? ur2 : exp_(k.vt_rai_x * dl_r)) : FT(0);   // CM1.terminal_velocity (Blk1M)
        const FT v_s = (q_sno > e) ? k.vt_sno_pref * (STD ? us2 : exp_(k.vt_sno_x * dl_s)) : FT(0);
        const FT dv = v_s - v_r;                                        // IEEE, reference order (cancellation)
        const FT dv_eff = sqrtg_(dv * dv + k.coeff_disp * (v_s * v_s + v_r * v_r));
        const FT common = inv_rho * n0_s * p.rain.n0 * dv_eff;
        // arm (i, j) = (snow, rain):  λ_i³λ_j^(δ+1) ... with δ = rain me+Δm
        const FT dr = k.rs_rai_delta, ds = k.rs_sno_delta;
        const FT pj_r = STD ? k.r0_rai4 * ur16 : exp_((dr + FT(1)) * ll_r);   // λ_r^(δr+1)
        const FT pj_s = STD ? k.r0_sno3 * us24 : exp_((ds + FT(1)) * ll_s);   // λ_s^(δs+1)
        const FT S_rai_sno = common * k.rs_rai_pref * (lam_s * pj_r) *
                             (FT(2) * (lam_s * lam_s) + k.rs_r_c1 * (lam_s * lam_r) + k.rs_r_c2 * (lam_r * lam_r));
        const FT S_sno_rai = common * k.rs_sno_pref * (lam_r * pj_s) *
                             (FT(2) * (lam_r * lam_r) + k.rs_s_c1 * (lam_r * lam_s) + k.rs_s_c2 * (lam_s * lam_s));
        const bool both = (q_rai > e) && (q_sno > e);
        const FT a = both ? S_rai_sno : FT(0);
        const FT b = both ? S_sno_rai : FT(0);
        r.s[S1M_ACCR_RAI_SNO_COLD] = is_warm ? FT(0) : a;
        r.s[S1M_ACCR_RAI_SNO_WARM] = is_warm ? b : FT(0);
        r.s[S1M_ACCR_MELT_RAI_SNO] = is_warm ? alpha_melt * a : FT(0);
    } else {
        r.s[S1M_ACCR_RAI_SNO_COLD] = r.s[S1M_ACCR_RAI_SNO_WARM] = r.s[S1M_ACCR_MELT_RAI_SNO] = FT(0);
    }

    // ---- ventilated diffusional growth / melt                           CM1:915-1139
    // a + b cbrt(Sc) Γvent sqrt(2 v0 χv/ν λ⁻¹) (λ⁻¹/r0)^((ve+Δv)/2)
    const FT vent_s = fma_(k.vent_sno_b, STD ? k.sqrt_r0_sno * us5 : exp_(fma_(k.vent_sno_x, dl_s, FT(0.5) * ll_s)), k.vent_sno_a);
    if (o.rain_condensation_evaporation) {
        const FT vent_r = fma_(k.vent_rai_b * sqrtg_(v0_r), STD ? k.sqrt_r0_rai * ur3 : exp_(fma_(k.vent_rai_x, dl_r, FT(0.5) * ll_r)), k.vent_rai_a);
        const FT rate = k.four_pi * p.rain.n0 * inv_rho * S_liq * G_liq * (lam_r * lam_r) * vent_r;
        r.s[S1M_PHASE_VAP_RAI] = cap0_((q_rai > e && S_liq < FT(0)) ? rate : FT(0));
    } else
        r.s[S1M_PHASE_VAP_RAI] = FT(0);
    if (o.snow_deposition_sublimation) {
        const FT rate = k.four_pi * n0_s * inv_rho * S_ice * G_ice * (lam_s * lam_s) * vent_s;
        const FT v = (q_sno > e) ? rate : FT(0);
        r.s[S1M_PHASE_VAP_SNO] = (o.snow_deposition_sublimation == CUMICRO_1M_SNOW_SUBLIMATION_ONLY) ? cap0_(v) : v;
    } else
        r.s[S1M_PHASE_VAP_SNO] = FT(0);
    const FT melt_common = k.four_pi * inv_rho * p.aps.K_therm * rcp_(Lf) * (T - T_freeze);
    r.s[S1M_MELT_ICL_LCL] = (o.cloud_ice_melt && q_icl > e && T > T_freeze) ? melt_common * p.cloud_ice.n0 * (lam_i * lam_i) : FT(0);
    r.s[S1M_MELT_SNO_RAI] = (o.snow_melt && q_sno > e && T > T_freeze) ? melt_common * n0_s * (lam_s * lam_s) * vent_s : FT(0);
    return r;
}

// BMT._aggregate_tendencies                                              BMT:227-252
template <class FT> CM_DEV void aggregate_tendencies_1m(const Src1M<FT>& r, FT (&out)[4]) {
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

// BMT._linearize + _linearized_implicit_step                               BMT:269-465
// (IEEE divisions and the reference's operation order throughout: the 2x2 solves
// subtract nearly equal products.)
template <class FT, bool STD = false>
CM_DEV void linearized_implicit_step_1m(const typename P<FT>::params_1m& p, const ThermoK<FT>& tk, const OneMK<FT>& k, FT rho, FT T,
                                        FT q_tot, FT q_lcl, FT q_icl, FT q_rai, FT q_sno, FT inv_dt, FT (&out)[4]) {
    // the temperature-only thermodynamic state once for the source terms and for q_sat below (same functions: same bits)
    const ThermoShared<FT> th = thermo_shared(tk, T);
    const Src1M<FT> r = microphysics_source_terms_1m<FT, STD>(p, tk, k, rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno, &th);
    const FT* s = r.s;
    const FT q_min = tk.q_min;
    const FT d_lcl = fmax_(q_min, q_lcl), d_icl = fmax_(q_min, q_icl), d_rai = fmax_(q_min, q_rai), d_sno = fmax_(q_min, q_sno);
    // 19 quotients over 4 divisors: one correctly rounded reciprocal per divisor (cm_math.cuh, divr_).  q_min = 0 (legal, not
    // sensible) makes the divisor of an empty species zero: the reference's 0/0 = NaN is NaN here as well (rcp of 0 is not finite).
    const FT r_lcl = rcp_cr_(d_lcl), r_icl = rcp_cr_(d_icl), r_rai = rcp_cr_(d_rai), r_sno = rcp_cr_(d_sno);
    auto over = [](FT x, FT d, FT r) { return divr_(x, d, r); };
    FT M11 = 0, M12 = 0, M22 = 0, M31 = 0, M33 = 0, M34 = 0, M41 = 0, M42 = 0, M43 = 0, M44 = 0, e1 = 0, e2 = 0, e4 = 0;
    FT D;
    bool src;
    D = over(s[S1M_PHASE_VAP_LCL], d_lcl, r_lcl); src = s[S1M_PHASE_VAP_LCL] >= FT(0);
    e1 += src ? s[S1M_PHASE_VAP_LCL] : FT(0); M11 += src ? FT(0) : D;
    D = over(s[S1M_PHASE_VAP_ICL], d_icl, r_icl); src = s[S1M_PHASE_VAP_ICL] >= FT(0);
    e2 += src ? s[S1M_PHASE_VAP_ICL] : FT(0); M22 += src ? FT(0) : D;
    D = over(s[S1M_MELT_ICL_LCL], d_icl, r_icl); M22 -= D; M12 += D;
    D = over(s[S1M_ACNV_LCL_RAI], d_lcl, r_lcl); M11 -= D; M31 += D;
    D = over(s[S1M_ACNV_ICL_SNO], d_icl, r_icl); M22 -= D; M42 += D;
    D = over(s[S1M_ACCR_LCL_RAI], d_lcl, r_lcl); M11 -= D; M31 += D;
    const FT D_cold = over(s[S1M_ACCR_LCL_SNO_COLD], d_lcl, r_lcl);
    const FT D_warm = over(s[S1M_ACCR_LCL_SNO_WARM], d_lcl, r_lcl);
    M11 -= D_cold + D_warm; M31 += D_warm; M41 += D_cold;
    D = over(s[S1M_ACCR_MELT_LCL_SNO], d_sno, r_sno); M44 -= D; M34 += D;
    D = over(s[S1M_ACCR_ICL_RAI], d_icl, r_icl); M22 -= D; M42 += D;
    D = over(s[S1M_ACCR_ICL_SNO], d_icl, r_icl); M22 -= D; M42 += D;
    D = over(s[S1M_ACCR_FREEZE_ICL_RAI], d_rai, r_rai); M33 -= D; M43 += D;
    D = over(s[S1M_ACCR_RAI_SNO_WARM], d_sno, r_sno); M44 -= D; M34 += D;
    D = over(s[S1M_ACCR_MELT_RAI_SNO], d_sno, r_sno); M44 -= D; M34 += D;
    D = over(s[S1M_ACCR_RAI_SNO_COLD], d_rai, r_rai); M33 -= D; M43 += D;
    D = over(-s[S1M_PHASE_VAP_RAI], d_rai, r_rai); M33 -= D;
    D = over(s[S1M_PHASE_VAP_SNO], d_sno, r_sno); src = s[S1M_PHASE_VAP_SNO] >= FT(0);
    e4 += src ? s[S1M_PHASE_VAP_SNO] : FT(0); M44 += src ? FT(0) : D;
    D = over(s[S1M_MELT_SNO_RAI], d_sno, r_sno); M44 -= D; M34 += D;

    // inv_dt = RN(1/dt), from the host (uniform)
    // q_sat over liquid / ice at the (unclamped) state                       BMT:409-412
    const FT rRT = rho * tk.R_v * T;
    const FT r_rRT = rcp_cr_(rRT);   // one correctly rounded reciprocal, two IEEE quotients (divr_, cm_math.cuh)
    const FT q_sat_min = fmin_(divr_(th.p_vs_l, rRT, r_rRT), divr_(th.p_vs_i, rRT, r_rRT));
    const FT q_v = q_tot - q_lcl - q_icl - q_rai - q_sno;
    const FT e_sum = fmax_(e1 + e2 + e4, tk.eps);   // >= eps > 0; the numerator is an exact zero wherever the air is subsaturated
    const FT alpha = fmin_(FT(1), divr_(clamp0_(q_v - q_sat_min) * inv_dt, e_sum, rcp_cr_(e_sum)));
    const FT a11 = inv_dt - M11, a12 = -M12, a22 = inv_dt - M22, a31 = -M31, a33 = inv_dt - M33, a34 = -M34, a41 = -M41, a42 = -M42,
             a43 = -M43, a44 = inv_dt - M44;
    const FT b1 = alpha * e1 + inv_dt * q_lcl;
    const FT b2 = alpha * e2 + inv_dt * q_icl;
    const FT b3 = inv_dt * q_rai;
    const FT b4 = alpha * e4 + inv_dt * q_sno;
    // det12 >= 1/dt² > 0 and det > 0 (BMT:449-450): positive normal divisors, two quotients each
    const FT det12 = a11 * a22;
    const FT r_det12 = rcp_cr_(det12);
    const FT q_lcl_new = divr_(b1 * a22 - a12 * b2, det12, r_det12);
    const FT q_icl_new = divr_(a11 * b2, det12, r_det12);
    const FT r3 = fma_(-a31, q_lcl_new, b3);
    const FT r4 = fma_(-a41, q_lcl_new, fma_(-a42, q_icl_new, b4));
    const FT det = fma_(-a34, a43, a33 * a44);
    const FT r_det = rcp_cr_(det);
    const FT q_rai_new = divr_(r3 * a44 - a34 * r4, det, r_det);
    const FT q_sno_new = divr_(a33 * r4 - r3 * a43, det, r_det);
    out[0] = (q_lcl_new - q_lcl) * inv_dt;
    out[1] = (q_icl_new - q_icl) * inv_dt;
    out[2] = (q_rai_new - q_rai) * inv_dt;
    out[3] = (q_sno_new - q_sno) * inv_dt;
}

// Δt, Δt/nsub and their IEEE reciprocals (host side: uniform over the grid)
template <class FT> struct LinAvgK { FT dt, inv_dt, dt_sub, inv_dt_sub; };
template <class FT> __host__ inline LinAvgK<FT> make_linavg_k(FT dt, int nsub) {
    LinAvgK<FT> lk;
    lk.dt = dt; lk.inv_dt = FT(1) / dt; lk.dt_sub = dt / FT(nsub); lk.inv_dt_sub = FT(1) / lk.dt_sub;
    return lk;
}

// BMT.bulk_microphysics_tendencies(::LinearizedAverage, ...)                 BMT:572-632
template <class FT, bool STD = false>
CM_DEV void bmt1m_linearized_average(const typename P<FT>::params_1m& p, const ThermoK<FT>& tk, const OneMK<FT>& k, FT rho, FT T,
                                     FT q_tot, FT q_lcl, FT q_icl, FT q_rai, FT q_sno, const LinAvgK<FT>& lk, int nsub, FT Lv_over_cp,
                                     FT Ls_over_cp, FT (&out)[4]) {
    const FT q0[4] = {q_lcl, q_icl, q_rai, q_sno};
    const FT dt_sub = lk.dt_sub;
    for (int it = 0; it < nsub; ++it) {
        FT rt[4];
        linearized_implicit_step_1m<FT, STD>(p, tk, k, rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno, lk.inv_dt_sub, rt);
        q_lcl += rt[0] * dt_sub;
        q_icl += rt[1] * dt_sub;
        q_rai += rt[2] * dt_sub;
        q_sno += rt[3] * dt_sub;
        T += (Lv_over_cp * (rt[0] + rt[2]) + Ls_over_cp * (rt[1] + rt[3])) * dt_sub;
    }
    out[0] = divr_(q_lcl - q0[0], lk.dt, lk.inv_dt);   // differences are exact zeros where nothing happens
    out[1] = divr_(q_icl - q0[1], lk.dt, lk.inv_dt);
    out[2] = divr_(q_rai - q0[2], lk.dt, lk.inv_dt);
    out[3] = divr_(q_sno - q0[3], lk.dt, lk.inv_dt);
}

// ---- terminal velocities of the 1-moment / non-equilibrium schemes -------------------------------
// CO.Chen2022_vel_coeffs(::Chen2022VelTypeSmallIce | ::Chen2022VelTypeLargeIce, ρₐ, ρᵢ) (CO:302-349):
// everything that depends on ρᵢ only is parameter-only and folded on the host; per point
// remain ρₐ^A, the bi(ρₐ) and 1000^bi.
template <class FT> struct ChenIceK {
    FT As, Es, Fs, Bs, Cs, Gs;                 // small ice (Table B3 -> B2)
    FT Al, Bl, Cl, El, Fl, Gl, Hl;             // large ice (Table B5 -> B4)
};

// CM1.terminal_velocity(precip, ::Blk1MVelType, ρ, q) with per-point λ⁻¹          CM1:223-249
template <class FT> CM_DEV FT terminal_velocity_blk1m(FT e, FT pref_v0, FT x, const MPSpeciesK<FT>& sk, FT log_rho_q, FT log_n0, FT q) {
    FT lam, ll;
    lambda_inverse(sk, log_rho_q, log_n0, lam, ll);
    const FT w = pref_v0 * exp_(x * (ll - sk.log_r0));
    return (q > e) ? w : FT(0);
}

// Σ_k CO.Chen2022_exponential_pdf(aiu_k, bi_k, ciu_k, λ⁻¹, 3)                       CO:414-422
template <class FT, int N> CM_DEV FT chen_exponential_pdf_sum3(const FT (&aiu)[N], const FT (&bi)[N], const FT (&ciu)[N], FT lam_inv) {
    const FT ll = log_full_(lam_inv), il = FT(1) / lam_inv;
    FT w = FT(0);
#pragma unroll
    for (int i = 0; i < N; ++i)
        w += aiu[i] * exp_full_(FT(-4) * ll - (bi[i] + FT(4)) * log_full_(il + ciu[i])) * tgamma_(bi[i] + FT(4)) / FT(6);
    return w;
}

}  // namespace cm
