// cm_sb2006.cuh — Seifert & Beheng (2006) 2-moment warm-rain processes and the
// non-equilibrium condensation/evaporation relaxation, fused per grid point.
//
// Device form of BMT.warm_rain_tendencies_2m (BMT:707-782) and its callees:
//   NEQ._conv_q_vap_to_q_lcl_const  NEQ:117-140     CM2.rain_evaporation  CM2:780-828
//   CM2.autoconversion              CM2:396-427     CM2.accretion         CM2:445-470
//   CM2.cloud_liquid_self_collection CM2:488-501    CM2.rain_self_collection CM2:545-560
//   CM2.rain_breakup                CM2:579-601     CM2.number_tendency_from_mass_limits CM2:882-891
//   CM2.pdf_rain_parameters         CM2:67-110      CM2.Γ_incl            CM2:746-753
// The reference evaluates pdf_rain_parameters three times and the saturation
// vapour pressure three times per point with identical arguments; here every
// shared quantity is computed once.  Regime predicates use the same comparison
// operators (<, <=, >=) on the same quantities as the reference.
#pragma once
#include "cm_thermo.cuh"

namespace cm {

// Per-launch constants derived on the host from the SB2006 parameter block.
template <class FT> struct SB2006K {
    FT acnv_pref;     // kcc/20/x_star (nu_c+2)(nu_c+4)/(nu_c+1)^2 rho0
    FT inv_x_star;    // 1/acnv.x_star
    FT lclsc_pref;    // kcc (nu_c+2)/(nu_c+1) rho0
    FT pi_rho_w;      // pi rho_w (rain pdf)
    FT cbrt_pi_rho_w;
    FT six_over_pi_rho_w;
    FT six_x_star_r;  // 6 * pdf_r.xr_min (evaporation t_star)
    FT inv_xr_min;    // 1/pdf_r.xr_min
    FT cbrt_Sc;       // cbrt(nu_air / max(D_vapor, eps))
    FT inv_nu_air;
    FT inv_K_safe, inv_D_safe;
    FT gi_c1[2], gi_e1[2], gi_c2[2], gi_de[2];  // Γ_incl coefficients for a = -1 and a = beta_vent_0 (de = e2 - e1)
    FT cbrt_six_over_pi_rho_w, cbrt_six_x_star_r, log_six_x_star_r_third, cbrt_one_sixth;
    FT inv_numadj_tau;
    FT inv_xc_min, inv_xc_max, inv_xr_max;
    FT two_pi;
    int pw_acnv_b, pw_accr_c, pw_self_d;   // pow_param_code of the three parameter exponents
    int same_rho0_evap, same_rho0_accr;    // evap.rho0 / accr.rho0 equal pdf_r.rho0 (one sqrt(rho0/rho) serves all)
};

template <class FT>
__host__ inline SB2006K<FT> make_sb2006_k(const typename P<FT>::sb2006& sb, const typename P<FT>::air& aps,
                                          bool method_is_f32 = false) {
    SB2006K<FT> k;
    const FT pi = FT(3.141592653589793238462643383279502884L);
    const FT nu_c = sb.pdf_c.nu_c;
    k.acnv_pref = sb.acnv.kcc / 20 / sb.acnv.x_star * (nu_c + 2) * (nu_c + 4) / ((nu_c + 1) * (nu_c + 1)) * sb.acnv.rho0;
    k.inv_x_star = FT(1) / sb.acnv.x_star;
    k.lclsc_pref = sb.acnv.kcc * (nu_c + 2) / (nu_c + 1) * sb.acnv.rho0;
    k.pi_rho_w = pi * sb.pdf_r.rho_w;
    k.cbrt_pi_rho_w = std::cbrt(k.pi_rho_w);
    k.six_over_pi_rho_w = FT(6) / (pi * sb.pdf_r.rho_w);
    k.six_x_star_r = FT(6) * sb.pdf_r.xr_min;
    k.inv_xr_min = FT(1) / sb.pdf_r.xr_min;
    const FT epsn = method_is_f32 ? FT(2.2737367544323206e-13) : FT(2.8126442852362996e-103);
    k.cbrt_Sc = std::cbrt(aps.nu_air / std::max(aps.D_vapor, epsn));
    k.inv_nu_air = FT(1) / aps.nu_air;
    k.inv_K_safe = FT(1) / std::max(aps.K_therm, epsn);
    k.inv_D_safe = FT(1) / std::max(aps.D_vapor, epsn);
    const FT a[2] = {FT(-1), sb.evap.beta_vent_0};
    for (int i = 0; i < 2; ++i) {
        k.gi_c1[i] = FT(0.33) - FT(0.7) * a[i];
        k.gi_e1[i] = FT(0.08) - FT(0.93) * a[i];
        k.gi_c2[i] = FT(1.34) - FT(0.1) * a[i];
        k.gi_de[i] = (FT(0.8) - a[i]) - k.gi_e1[i];
    }
    k.cbrt_six_over_pi_rho_w = std::cbrt(k.six_over_pi_rho_w);
    k.cbrt_six_x_star_r = std::cbrt(k.six_x_star_r);
    k.log_six_x_star_r_third = std::log(k.six_x_star_r) / 3;
    k.cbrt_one_sixth = std::cbrt(FT(1) / FT(6));
    k.inv_numadj_tau = FT(1) / sb.numadj_tau;
    k.inv_xc_min = FT(1) / sb.pdf_c.xc_min;
    k.inv_xc_max = FT(1) / sb.pdf_c.xc_max;
    k.inv_xr_max = FT(1) / sb.pdf_r.xr_max;
    k.two_pi = 2 * pi;
    k.pw_acnv_b = pow_param_code(sb.acnv.b); k.pw_accr_c = pow_param_code(sb.accr.c); k.pw_self_d = pow_param_code(sb.self.d);
    k.same_rho0_evap = sb.evap.rho0 == sb.pdf_r.rho0; k.same_rho0_accr = sb.accr.rho0 == sb.pdf_r.rho0;
    return k;
}

// SPEC of a parameter block (see warm_rain_tendencies_2m): 0 / 1 = default structure with the not-limited / limited rain PSD, -1 = generic
template <class FT> __host__ inline int sb2006_spec(const typename P<FT>::sb2006& sb) {
    const bool std_structure = sb.acnv.b == FT(3) && sb.accr.c == FT(4) && sb.self.d == FT(-5) && sb.evap.rho0 == sb.pdf_r.rho0 &&
                               sb.accr.rho0 == sb.pdf_r.rho0;
    return std_structure ? (sb.pdf_r.limited ? 1 : 0) : -1;
}

template <class FT> struct RainPDF { FT N0r, Dr_mean, xr_mean, lam; };

// CM2.pdf_rain_parameters                                          CM2:67-110
// (q and N are the caller's already-floored safe values, as at every reference call site.)
// LIM: -1 = the variant is read from the block at run time, 0 / 1 = known at compile time (SB2006Spec below).
// cbrt_pi_rho_w: cbrt(π ρw) from the host (SB2006K) or 0 = not supplied (Eq. 95 then takes the cube root of the quotient).
template <class FT, int LIM = -1, bool HAVE_CBRT = false>
CM_DEV RainPDF<FT> pdf_rain_parameters(const typename P<FT>::sb_pdf_r& pdf, FT pi_rho_w, FT e, FT q, FT rho, FT N,
                                       FT cbrt_pi_rho_w = FT(0)) {
    const FT safe_q = fmax_(q, e);
    const FT safe_N = fmax_(N, e);
    const FT L = rho * safe_q;
    RainPDF<FT> r;
    const bool limited = (LIM < 0) ? (pdf.limited != 0) : (LIM == 1);
    if (!limited) {
        const FT xr_mean = L * rcp_(safe_N);
        const FT lam = cbrtp_(pi_rho_w * safe_N * rcp_(L));
        const bool cond = (N < e) || (q < e);
        r.lam = lam;
        r.N0r = cond ? FT(0) : lam * safe_N;
        r.Dr_mean = cond ? FT(0) : rcp_(lam);
        r.xr_mean = cond ? FT(0) : xr_mean;
    } else {
        const FT inv_L = rcp_(L);
        const FT xt = clamp_(L * rcp_(safe_N), pdf.xr_min, pdf.xr_max);                       // SB2006 Eq. (94)
        const FT c95 = HAVE_CBRT ? cbrt_pi_rho_w * rcbrtp_(xt) : cbrtp_(pi_rho_w * rcp_(xt));   // cbrt(π ρw / xt)
        const FT N0r = clamp_(safe_N * c95, pdf.N0_min, pdf.N0_max);  // Eq. (95)
        const FT lam = clamp_(sqrtp_(sqrtp_(pi_rho_w * N0r * inv_L)), pdf.lam_min, pdf.lam_max);  // Eq. (96)
        const FT xr_mean = clamp_(L * lam * rcp_(N0r), pdf.xr_min, pdf.xr_max);                // Eq. (97)
        const bool cond = (N < e) && (q < e);
        r.lam = lam;
        r.N0r = cond ? FT(0) : N0r;
        r.Dr_mean = cond ? FT(0) : rcp_(lam);
        r.xr_mean = cond ? FT(0) : xr_mean;
    }
    return r;
}

// CM2.number_tendency_from_mass_limits                              CM2:882-891
template <class FT> CM_DEV FT number_tendency_from_mass_limits(FT e, FT inv_x_min, FT inv_x_max, FT inv_tau, FT q, FT n) {
    const FT n_target = (q < e) ? FT(0) : clamp_(n, q * inv_x_max, q * inv_x_min);
    return (n_target - n) * inv_tau;
}

template <class FT> struct Warm2M {
    FT dq_lcl_dt, dn_lcl_dt, dq_rai_dt, dn_rai_dt;
    FT leaf[CUMICRO_SB2006_NLEAF];
};

// BMT.bulk_microphysics_tendencies(::Microphysics2Moment, mp{WR,Nothing}, ...)  BMT:820-854
// q_ice is the cloud-ice content seen by the thermodynamics (0 for warm-only).
//
// Where the reference subtracts nearly equal numbers (tau = 1 - q_l/(q_l+q_r), and the
// (1 - tau) it forms from it) the operations are IEEE and in the reference's order, so
// the rounding pattern is the reference's own.
//
// SPEC (compile-time specialisation for the structure of the reference's default SB2006 block): SPEC < 0 = generic; otherwise
// bit 0 = the rain PSD variant (limited), and the parameter structure is the default one — exponents acnv.b = 3, accr.c = 4,
// self.d = -5 and evap.rho0 = accr.rho0 = pdf_r.rho0 (host-checked, sb2006_spec()).  The uniform run-time branches on these were
// if-converted by the compiler: both IEEE square roots of the "different rho0" arms and the switch of pow_param executed at
// every point (ncu source view: 3.6 % + 3.8 % of the kernel's instructions).
template <class FT, int SPEC = -1>
CM_DEV Warm2M<FT> warm_rain_tendencies_2m(const typename P<FT>::params_2m_warm& p, const ThermoK<FT>& tk,
                                          const SB2006K<FT>& sk, FT rho, FT T, FT q_tot, FT q_lcl, FT n_lcl,
                                          FT q_rai, FT n_rai, FT q_ice, FT N_rai_given = FT(-1)) {
    const FT e = tk.eps;
    const auto& sb = p.sb;
    Warm2M<FT> o;

    // input clamps                                                   BMT:827-836
    rho = clamp0_(rho);
    q_tot = clamp0_(q_tot);
    q_lcl = clamp0_(q_lcl);
    q_rai = clamp0_(q_rai);
    n_lcl = clamp0_(n_lcl);
    n_rai = clamp0_(n_rai);
    const FT N_lcl = rho * n_lcl;   // BMT:718-719
    // (the stand-alone leaf CM2.rain_evaporation takes the number density N_rai itself: N_rai_given >= 0)
    const FT N_rai = (N_rai_given >= FT(0)) ? N_rai_given : rho * n_rai;
    const FT inv_rho = rcp_(rho);

    // ---- thermodynamic state shared by cond/evap and rain evaporation
    const TempState<FT> ts = temp_state(tk, T);
    const FT p_vs = p_sat_liq(tk, ts);
    const FT inv_p_vs = rcp_(fmax_(p_vs, tk.eps_n));
    const FT Lv = latent_heat_vapor(tk, T);
    const FT q_liq = q_lcl + q_rai;
    const FT qv = q_vap(q_tot, q_liq, q_ice);
    const FT rho_Rv_T = rho * tk.R_v * T;
    const FT qv_sat = p_vs * rcp_(rho_Rv_T);
    const FT sat_excess = qv - qv_sat;

    // ---- NEQ._conv_q_vap_to_q_lcl_const                            NEQ:117-140
    {
        const FT cp_air = cp_m(tk, q_tot, q_liq, q_ice);
        const FT dqsl_dT = qv_sat * fma_(Lv * tk.inv_R_v * ts.inv_T, ts.inv_T, -ts.inv_T);  // NEQ.dqcld_dT
        // 1/(tau Gamma), Gamma = 1 + Lv/cp_air dqsl_dT                                     NEQ.gamma_helper
        const FT inv_ts = cp_air * rcp_(p.condevap_tau_relax * fma_(Lv, dqsl_dT, cp_air));
        const FT cond = (sat_excess < FT(0)) ? -fmin_(-sat_excess, q_lcl) * inv_ts : sat_excess * inv_ts;
        o.leaf[CUMICRO_SB_COND_DQ_LCL] = cond;
    }

    // ---- rain size distribution (shared by evaporation, self-collection, breakup)
    const FT safe_q_rai = fmax_(q_rai, e);
    const FT safe_N_rai = fmax_(N_rai, e);
    constexpr bool STD = SPEC >= 0;
    const RainPDF<FT> rp = pdf_rain_parameters<FT, STD ? (SPEC & 1) : -1, true>(sb.pdf_r, sk.pi_rho_w, e, safe_q_rai, rho, safe_N_rai,
                                                                               sk.cbrt_pi_rho_w);
    const FT xr_mean = rp.xr_mean;
    // every power of xr_mean below comes from ONE cube root and ONE logarithm
    FT inv_cx;                                  // xr_mean^(-1/3) comes out of the same iteration as the cube root
    const FT cx = cbrt_pair_(xr_mean, inv_cx);
    const FT inv_xr_mean = inv_cx * inv_cx * inv_cx;
    const FT Dr = cx * sk.cbrt_six_over_pi_rho_w;  // cbrt(6 xr/(pi rho_w)): mean-volume diameter  CM2:590, 802
    const FT sqrt_rho0_rho = sqrtp_(sb.pdf_r.rho0 * inv_rho);
    const bool no_rain = (q_rai < e) || (N_rai < e);

    // ---- CM2.rain_evaporation                                       CM2:780-828
    {
        const FT S = fma_(qv * rho_Rv_T, inv_p_vs, FT(-1));                  // TDI.supersaturation_over_liquid
        const FT G = G_func(tk, sk.inv_K_safe, sk.inv_D_safe, Lv, inv_p_vs, ts);  // CO.G_func_liquid
        const FT lx = logp_(xr_mean);
        const FT t_star = sk.cbrt_six_x_star_r * inv_cx;           // cbrt(6 x*/xr)
        const FT lt = fma_(lx, FT(-1.0 / 3.0), sk.log_six_x_star_r_third);  // log(t_star)
        // Γ_incl(a, t) = exp(-t) / (c1 t^e1 + c2 t^e2) = exp(-t - e1 ln t) / (c1 + c2 t^(e2-e1))   CM2:746-753
        // The not-limited PSD leaves xr_mean unclamped: t_star = cbrt(6 x*/xr_mean) exceeds exp_'s |x| <= 708 domain for tiny q_rai with
        // leftover n_rai (the reference's exp underflows to 0 there; exp_ would wrap its exponent).  The limited PSD bounds t_star by
        // cbrt(6).  `lim` is uniform over the grid: one branch, no divergence.
        const bool lim = STD ? ((SPEC & 1) != 0) : (sb.pdf_r.limited != 0);
        const FT ga0 = fma_(-sk.gi_e1[0], lt, -t_star), ga1 = fma_(-sk.gi_e1[1], lt, -t_star);
        const FT ge0 = lim ? exp_(ga0) : exp_full_(ga0), ge1 = lim ? exp_(ga1) : exp_full_(ga1);
        const FT gd0 = sk.gi_de[0] * lt, gd1 = sk.gi_de[1] * lt;
        const FT gx0 = lim ? exp_(gd0) : exp_full_(gd0), gx1 = lim ? exp_(gd1) : exp_full_(gd1);
        const FT gi0 = ge0 * rcp_(fma_(sk.gi_c2[0], gx0, sk.gi_c1[0]));
        const FT gi1 = ge1 * rcp_(fma_(sk.gi_c2[1], gx1, sk.gi_c1[1]));
        const FT a_vent_0 = sb.evap.a_vent_0_coeff * gi0;
        const FT b_vent_0 = sb.evap.b_vent_0_coeff * gi1;
        FT sqrt_rho0e = sqrt_rho0_rho;
        if (!STD && !sk.same_rho0_evap) sqrt_rho0e = sqrt_(sb.evap.rho0 * inv_rho);
        const FT N_Re = sb.evap.alpha * exp_(sb.evap.beta * lx) * sqrt_rho0e * Dr * sk.inv_nu_air;
        const FT v = sk.cbrt_Sc * sqrtp_(N_Re);
        const FT Fv0 = fma_(b_vent_0, v, a_vent_0);
        const FT Fv1 = fma_(sb.evap.b_vent_1, v, sb.evap.a_vent_1);
        const FT common = sk.two_pi * G * S * N_rai * Dr;
        const FT dn = cap0_(common * Fv0 * inv_xr_mean);
        const FT dq = cap0_(common * Fv1 * inv_rho);
        const bool off_q = (q_rai < e) || (N_rai <= e) || (S >= FT(0));
        const bool off_n = off_q || (xr_mean * sk.inv_xr_min < e);
        o.leaf[CUMICRO_SB_EVAP_DN_RAI] = off_n ? FT(0) : dn;
        o.leaf[CUMICRO_SB_EVAP_DQ_RAI] = off_q ? FT(0) : dq;
    }

    // ---- CM2.autoconversion + CM2.accretion (shared tau)             CM2:396-470
    {
        const FT safe_q_lcl = fmax_(q_lcl, e);
        const FT safe_N_lcl = fmax_(N_lcl, e);
        const FT L_lcl = rho * safe_q_lcl;
        const FT inv_L_lcl = rcp_(L_lcl);
        const FT x_lcl = fmin_(sb.acnv.x_star, L_lcl * rcp_(safe_N_lcl));
        const FT tau = FT(1) - div_(safe_q_lcl, safe_q_lcl + q_rai);          // SB2006 Eq. (5), IEEE, reference order
        const FT one_m_tau = FT(1) - tau;
        const FT tau_a = powp_(tau, sb.acnv.a);
        const FT phi_au = (q_rai < e) ? FT(0) : sb.acnv.A * tau_a * (STD ? pow_int_<3>(FT(1) - tau_a) : pow_param(FT(1) - tau_a, sb.acnv.b, sk.pw_acnv_b));
        const FT LL = L_lcl * L_lcl;
        const FT dL_rai_dt = sk.acnv_pref * LL * (x_lcl * x_lcl) *
                             fma_(phi_au, rcp_(one_m_tau * one_m_tau), FT(1)) * inv_rho;   // Eq. (4)
        const FT dN_rai_dt = dL_rai_dt * sk.inv_x_star;
        const bool off = (q_lcl < e) || (N_lcl < e);
        const FT dq = off ? FT(0) : dL_rai_dt * inv_rho;
        const FT dN_rai = off ? FT(0) : dN_rai_dt;
        const FT dN_lcl_au = off ? FT(0) : FT(-2) * dN_rai_dt;
        o.leaf[CUMICRO_SB_ACNV_DQ_LCL] = -dq;
        o.leaf[CUMICRO_SB_ACNV_DN_LCL] = dN_lcl_au;
        o.leaf[CUMICRO_SB_ACNV_DQ_RAI] = dq;
        o.leaf[CUMICRO_SB_ACNV_DN_RAI] = dN_rai;

        // CM2.cloud_liquid_self_collection (uses the unclamped q_lcl)   CM2:488-501
        const FT Lu = rho * q_lcl;
        const FT sc = -sk.lclsc_pref * inv_rho * (Lu * Lu) - dN_lcl_au;
        o.leaf[CUMICRO_SB_LCL_SELFCOL] = (q_lcl < e) ? FT(0) : sc;

        // CM2.accretion                                                  CM2:445-470
        const FT L_rai = rho * safe_q_rai;
        // (accretion floors q_rai at eps instead of 0, CM2:452, but is gated off below eps: same tau)
        FT sqrt_rho0a = sqrt_rho0_rho;
        if (!STD && !sk.same_rho0_accr) sqrt_rho0a = sqrt_(sb.accr.rho0 * inv_rho);
        const FT phi_arg = tau * rcp_(tau + sb.accr.tau0);
        const FT phi_ac = STD ? pow_int_<4>(phi_arg) : pow_param(phi_arg, sb.accr.c, sk.pw_accr_c);   // Eq. (8)
        const FT dLr = sb.accr.kcr * L_lcl * L_rai * phi_ac * sqrt_rho0a;             // Eq. (7)
        const bool off_ac = (q_lcl < e) || (q_rai < e) || (N_lcl < e);
        const FT dq_ac = off_ac ? FT(0) : dLr * inv_rho;
        o.leaf[CUMICRO_SB_ACCR_DQ_LCL] = -dq_ac;
        o.leaf[CUMICRO_SB_ACCR_DN_LCL] = off_ac ? FT(0) : -dLr * safe_N_lcl * inv_L_lcl;  // dL_lcl_dt / x_lcl
        o.leaf[CUMICRO_SB_ACCR_DQ_RAI] = dq_ac;
    }

    // ---- CM2.rain_self_collection / rain_breakup                       CM2:545-601
    {
        const FT L_rai = rho * safe_q_rai;
        const FT inv_Br = cx * sk.cbrt_one_sixth;   // 1/Br, Br = cbrt(6/xr_mean)   CM2:141-146
        const FT sc_arg = fma_(sb.self.kappa_rr, inv_Br, FT(1));
        FT sc = -sb.self.krr * N_rai * L_rai * sqrt_rho0_rho * (STD ? pow_int_<-5>(sc_arg) : pow_param(sc_arg, sb.self.d, sk.pw_self_d));
        sc = no_rain ? FT(0) : sc;
        const FT dD = Dr - sb.brek.Deq;
        const bool lim_b = STD ? ((SPEC & 1) != 0) : (sb.pdf_r.limited != 0);   // Dr is bounded only under the limited PSD
        const FT phi_p1 = (Dr < sb.brek.Dr_th) ? FT(0)
                                               : ((Dr <= sb.brek.Deq) ? fma_(sb.brek.kbr, dD, FT(1))
                                                                      : (lim_b ? exp_(sb.brek.kappa_br * dD) : exp_full_(sb.brek.kappa_br * dD)));
        const FT br = no_rain ? FT(0) : -phi_p1 * sc;   // Eq. (13): -(Φ_br + 1) dN_sc
        o.leaf[CUMICRO_SB_RAI_SELFCOL] = sc;
        o.leaf[CUMICRO_SB_RAI_BREAKUP] = br;
    }

    // ---- number adjustment (Horn 2012)                                 BMT:771-779
    o.leaf[CUMICRO_SB_NUMADJ_LCL] =
        number_tendency_from_mass_limits<FT>(e, sk.inv_xc_min, sk.inv_xc_max, sk.inv_numadj_tau, q_lcl, n_lcl);
    o.leaf[CUMICRO_SB_NUMADJ_RAI] =
        number_tendency_from_mass_limits<FT>(e, sk.inv_xr_min, sk.inv_xr_max, sk.inv_numadj_tau, q_rai, n_rai);

    // ---- aggregate in the order of BMT:736-779
    o.dq_lcl_dt = o.leaf[CUMICRO_SB_COND_DQ_LCL] + o.leaf[CUMICRO_SB_ACNV_DQ_LCL] + o.leaf[CUMICRO_SB_ACCR_DQ_LCL];
    o.dq_rai_dt = o.leaf[CUMICRO_SB_EVAP_DQ_RAI] + o.leaf[CUMICRO_SB_ACNV_DQ_RAI] + o.leaf[CUMICRO_SB_ACCR_DQ_RAI];
    o.dn_lcl_dt = fma_(o.leaf[CUMICRO_SB_ACNV_DN_LCL] + o.leaf[CUMICRO_SB_LCL_SELFCOL] + o.leaf[CUMICRO_SB_ACCR_DN_LCL],
                       inv_rho, o.leaf[CUMICRO_SB_NUMADJ_LCL]);
    o.dn_rai_dt = fma_(o.leaf[CUMICRO_SB_EVAP_DN_RAI] + o.leaf[CUMICRO_SB_ACNV_DN_RAI] + o.leaf[CUMICRO_SB_RAI_SELFCOL] +
                           o.leaf[CUMICRO_SB_RAI_BREAKUP],
                       inv_rho, o.leaf[CUMICRO_SB_NUMADJ_RAI]);
    return o;
}

}  // namespace cm

// ============================ 2-moment terminal velocities ============================
namespace cm {

// CM2.rain_terminal_velocity(::SB2006, ::SB2006VelType, q_rai, rho, N_rai)   CM2:685-702
// (+ _sb_rain_terminal_velocity_helper CM2:720-739: limited -> (1,1,1,1)).
template <class FT>
CM_DEV void rain_terminal_velocity_sb(const typename P<FT>::sb_pdf_r& pdf_r, const typename P<FT>::vel_sb2006& vel,
                                      FT pi_rho_w, FT e, FT q_rai, FT rho, FT N_rai, FT& vt0, FT& vt1) {
    const RainPDF<FT> r = pdf_rain_parameters<FT>(pdf_r, pi_rho_w, e, fmax_(q_rai, e), rho, fmax_(N_rai, e));
    const FT Dr_mean = r.Dr_mean;
    FT pa0 = FT(1), pb0 = FT(1), pa1 = FT(1), pb1 = FT(1);
    if (!pdf_r.limited) {
        const FT lam = FT(1) / Dr_mean;
        const FT two_rc = -log_full_(vel.aR / vel.bR) / vel.cR;  // 2 rc, rc = -1/(2 cR) log(aR/bR)
        const FT ta = two_rc * lam, tb = two_rc * (lam + vel.cR);
        const FT ea = exp_full_(-ta), eb = exp_full_(-tb);
        pa0 = ea;
        pb0 = eb;
        pa1 = (ta * ta * ta + FT(3) * (ta * ta) + FT(6) * ta + FT(6)) * ea / FT(6);
        pb1 = (tb * tb * tb + FT(3) * (tb * tb) + FT(6) * tb + FT(6)) * eb / FT(6);
    }
    const FT s = sqrt_(vel.rho0 / rho);
    const FT d1 = FT(1) + vel.cR * Dr_mean;
    const FT d2 = d1 * d1;
    const FT v0 = clamp0_(s * (vel.aR * pa0 - vel.bR * pb0 / d1));
    const FT v1 = clamp0_(s * (vel.aR * pa1 - vel.bR * pb1 / (d2 * d2)));
    vt0 = (N_rai < e) ? FT(0) : v0;
    vt1 = (q_rai < e) ? FT(0) : v1;
}

// CO.Chen2022_vel_coeffs(::Chen2022VelTypeRain, rho)                      CO:290-300
template <class FT>
CM_DEV void chen2022_vel_coeffs_rain(const typename P<FT>::vel_chen_rain& v, FT rho, FT aiu[3], FT bi[3], FT ciu[3]) {
    rho = fmax_(rho, FT(0));
    const FT q = exp_full_(v.rho0 * rho);
    const FT log1000 = FT(6.907755278982137);
    FT ai[3] = {v.a[0] * q, v.a[1] * q, v.a[2] * q * pow_full_(rho, v.a3_pow)};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        bi[i] = v.b[i] - v.b_rho * rho;
        aiu[i] = ai[i] * exp_full_(bi[i] * log1000);  // 1000^bi
        ciu[i] = v.c[i] * FT(1000);
    }
}

// CO.Chen2022_exponential_pdf(a, b, c, lam_inv, k), k! passed as inv_fac    CO:414-422
template <class FT> CM_DEV FT chen2022_exponential_pdf(FT a, FT b, FT c, FT log_lam_inv, FT inv_lam_inv, FT delta, FT inv_fac) {
    return a * exp_full_(-delta * log_lam_inv - (b + delta) * log_full_(inv_lam_inv + c)) * tgamma_(b + delta) * inv_fac;
}

// CM2.rain_terminal_velocity(::SB2006, ::Chen2022VelTypeRain, ...)          CM2:703-719
template <class FT>
CM_DEV void rain_terminal_velocity_chen(const typename P<FT>::sb_pdf_r& pdf_r, const typename P<FT>::vel_chen_rain& vel,
                                        FT pi_rho_w, FT e, FT q_rai, FT rho, FT N_rai, FT& vt0, FT& vt1) {
    FT aiu[3], bi[3], ciu[3];
    chen2022_vel_coeffs_rain<FT>(vel, rho, aiu, bi, ciu);
    const RainPDF<FT> r = pdf_rain_parameters<FT>(pdf_r, pi_rho_w, e, fmax_(q_rai, e), rho, fmax_(N_rai, e));
    const FT ll = log_full_(r.Dr_mean), il = FT(1) / r.Dr_mean;
    FT v0 = FT(0), v3 = FT(0);
#pragma unroll
    for (int i = 0; i < 3; ++i) v0 += chen2022_exponential_pdf<FT>(aiu[i], bi[i], ciu[i], ll, il, FT(1), FT(1));
#pragma unroll
    for (int i = 0; i < 3; ++i) v3 += chen2022_exponential_pdf<FT>(aiu[i], bi[i], ciu[i], ll, il, FT(4), FT(1.0 / 6.0));
    vt0 = (N_rai < e) ? FT(0) : clamp0_(v0);
    vt1 = (q_rai < e) ? FT(0) : clamp0_(v3);
}

// CM2.cloud_terminal_velocity                                               CM2:647-664
// with CM2.log_pdf_cloud_parameters_mass (CM2:176-190) and DT.generalized_gamma_Mn (DT:109-112).
// `gratio[2]` = Gamma((nu+1+n)/mu)/Gamma((nu+1)/mu) for n = 2/3, 5/3 and `pref0` =
// 1/18 cbrt((6/rho_w/pi)^2) grav/nu_air are parameter-only (host-side).
template <class FT>
CM_DEV void cloud_terminal_velocity(const typename P<FT>::sb_pdf_c& pdf_c, const typename P<FT>::vel_stokes& vel,
                                    FT pref0, const FT gratio[2], FT e, FT q_liq, FT rho, FT N_liq, FT& vt0, FT& vt1) {
    const FT safe_q = fmax_(q_liq, e), safe_N = fmax_(N_liq, e);
    const FT logx = log_full_(rho * safe_q / safe_N);
    const FT logB = -pdf_c.mu_c * (logx + pdf_c.loggamma_z1 - pdf_c.loggamma_z2);
    const FT pref = pref0 * (vel.rho_w / rho - FT(1));
    const FT inv_mu = FT(1) / pdf_c.mu_c;
    // M^n / N = B^(-n/mu) * gratio
    const FT m23 = exp_full_(-(FT(2.0 / 3.0) * inv_mu) * logB) * gratio[0];
    const FT m53 = exp_full_(-(FT(5.0 / 3.0) * inv_mu) * logB) * gratio[1];
    const FT v0 = pref * m23;
    const FT v1 = pref * (safe_N * m53) / rho / safe_q;
    const bool cond = (N_liq < e) || (q_liq < e);
    vt0 = cond ? FT(0) : v0;
    vt1 = cond ? FT(0) : v1;
}

}  // namespace cm
