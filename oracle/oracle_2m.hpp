// oracle_2m.hpp — CPU restatement of the Seifert-Beheng 2006 2-moment set and
// the non-equilibrium cond/evap relaxation.  TEST INFRASTRUCTURE ONLY (see
// oracle_base.hpp).  Follows src/Microphysics2M.jl, src/MicrophysicsNonEq.jl
// and BMT:707-854 of the reference line by line.
#pragma once
#include "oracle_base.hpp"

namespace orc {

// ---- MicrophysicsNonEq.jl -----------------------------------------------------
// NEQ.dqcld_dT                                                    NEQ:74-76
template <class FT> inline FT dqcld_dT(FT qv_sat, FT L, FT Rv, FT T) {
    return qv_sat * (L / (Rv * (T * T)) - 1 / T);
}
// NEQ.gamma_helper                                                NEQ:88-90
template <class FT> inline FT gamma_helper(FT L, FT cp_air, FT dq) { return 1 + (L / cp_air) * dq; }

// NEQ._conv_q_vap_to_q_lcl_const                                  NEQ:117-140
template <class FT>
inline FT conv_q_vap_to_q_lcl_const(FT tau, const Thermo<FT>& tps, FT q_tot, FT q_lcl, FT q_icl,
                                    FT q_rai, FT q_sno, FT rho, FT T) {
    FT Rv = tps.R_v();
    FT Lv = tps.L_v(T);
    FT cp_air = tps.cp_m(q_tot, q_lcl + q_rai, q_icl + q_sno);
    FT qv = Thermo<FT>::q_vap(q_tot, q_lcl + q_rai, q_icl + q_sno);
    FT qv_sat_liq = tps.q_sat_liq(T, rho);
    FT dqsl_dT = dqcld_dT(qv_sat_liq, Lv, Rv, T);
    FT Gam = gamma_helper(Lv, cp_air, dqsl_dT);
    FT sat_excess = qv - qv_sat_liq;
    FT timescale = tau * Gam;
    return (sat_excess < 0) ? -jmin(-sat_excess, jmax(FT(0), q_lcl)) / timescale
                            : sat_excess / timescale;
}

// NEQ._conv_q_vap_to_q_icl_const (+ INP_limiter NEQ:58-60)          NEQ:168-193
template <class FT>
inline FT conv_q_vap_to_q_icl_const(FT tau, const Thermo<FT>& tps, FT q_tot, FT q_lcl, FT q_icl,
                                    FT q_rai, FT q_sno, FT rho, FT T) {
    FT Rv = tps.R_v();
    FT Ls = tps.L_s(T);
    FT cp_air = tps.cp_m(q_tot, q_lcl + q_rai, q_icl + q_sno);
    FT qv = Thermo<FT>::q_vap(q_tot, q_lcl + q_rai, q_icl + q_sno);
    FT qv_sat_ice = tps.q_sat_ice(T, rho);
    FT dqsi_dT = dqcld_dT(qv_sat_ice, Ls, Rv, T);
    FT Gam = gamma_helper(Ls, cp_air, dqsi_dT);
    FT sat_excess = qv - qv_sat_ice;
    FT timescale = tau * Gam;
    FT tendency = (sat_excess < 0) ? -jmin(-sat_excess, jmax(FT(0), q_icl)) / timescale
                                   : sat_excess / timescale;
    bool limiter = (T > tps.T_freeze()) && (tendency > FT(0));
    return limiter ? FT(0) : tendency;
}

// ---- Microphysics2M.jl --------------------------------------------------------
template <class FT> struct RainPDF { FT N0r, Dr_mean, xr_mean; };

// CM2.pdf_rain_parameters (notlimited CM2:67-86 / limited CM2:87-110)
template <class FT>
inline RainPDF<FT> pdf_rain_parameters(const typename PT<FT>::sb_pdf_r& pdf, FT q, FT rho, FT N) {
    const FT eM = eps_2M<FT>(), eN = eps_2M<FT>();
    FT safe_q = jmax(q, eM);
    FT safe_N = jmax(N, eN);
    FT L = rho * safe_q;
    RainPDF<FT> r;
    if (!pdf.limited) {
        FT xr_mean = L / safe_N;
        FT lam = cbrt_(pi<FT>() * pdf.rho_w / xr_mean);
        FT N0r = lam * safe_N;
        FT Dr_mean = 1 / lam;
        bool cond = (N < eN) || (q < eM);
        r.N0r = cond ? FT(0) : N0r;
        r.Dr_mean = cond ? FT(0) : Dr_mean;
        r.xr_mean = cond ? FT(0) : xr_mean;
    } else {
        FT xt = jclamp(L / safe_N, pdf.xr_min, pdf.xr_max);                                 // Eq. (94)
        FT N0r = jclamp(safe_N * cbrt_(pi<FT>() * pdf.rho_w / xt), pdf.N0_min, pdf.N0_max);  // (95)
        FT lam = jclamp(sqrt_(sqrt_(pi<FT>() * pdf.rho_w * N0r / L)), pdf.lam_min, pdf.lam_max);  // (96)
        FT xr_mean = jclamp(L * lam / N0r, pdf.xr_min, pdf.xr_max);                         // (97)
        FT Dr_mean = 1 / lam;
        bool cond = (N < eN) && (q < eM);
        r.N0r = cond ? FT(0) : N0r;
        r.Dr_mean = cond ? FT(0) : Dr_mean;
        r.xr_mean = cond ? FT(0) : xr_mean;
    }
    return r;
}

// CM2.pdf_rain_parameters_mass                                     CM2:141-146
template <class FT>
inline void pdf_rain_parameters_mass(const typename PT<FT>::sb_pdf_r& pdf, FT q, FT rho, FT N,
                                     FT& Ar, FT& Br) {
    RainPDF<FT> r = pdf_rain_parameters<FT>(pdf, q, rho, N);
    Br = cbrt_(6 / r.xr_mean);
    Ar = N * Br / 3;
}

// CM2.log_pdf_cloud_parameters_mass                                CM2:176-190
template <class FT>
inline void log_pdf_cloud_parameters_mass(const typename PT<FT>::sb_pdf_c& pdf, FT q, FT rho, FT N,
                                          FT& logA, FT& logB) {
    const FT eM = eps_2M<FT>(), eN = eps_2M<FT>();
    FT safe_q = jmax(q, eM);
    FT safe_N = jmax(N, eN);
    FT L = rho * safe_q;
    FT logx = log_(L / safe_N);
    FT z1 = (pdf.nu_c + 1) / pdf.mu_c;
    FT lB = -pdf.mu_c * (logx + pdf.loggamma_z1 - pdf.loggamma_z2);
    FT lA = log_(pdf.mu_c) + log_(safe_N) + z1 * lB - pdf.loggamma_z1;
    bool cond = (N < eN) || (q < eM);
    logA = cond ? -inf<FT>() : lA;
    logB = cond ? inf<FT>() : lB;
}
// CM2.pdf_cloud_parameters_mass                                    CM2:199-202
template <class FT>
inline void pdf_cloud_parameters_mass(const typename PT<FT>::sb_pdf_c& pdf, FT q, FT rho, FT N,
                                      FT& Ac, FT& Bc) {
    FT lA, lB;
    log_pdf_cloud_parameters_mass<FT>(pdf, q, rho, N, lA, lB);
    Ac = exp_(lA);
    Bc = exp_(lB);
}

template <class FT> struct LclRaiRates { FT dq_lcl_dt, dN_lcl_dt, dq_rai_dt, dN_rai_dt; };

// CM2.autoconversion                                               CM2:396-427
template <class FT>
inline LclRaiRates<FT> autoconversion(const typename PT<FT>::sb_acnv& acnv,
                                      const typename PT<FT>::sb_pdf_c& pdf_c, FT q_lcl, FT q_rai,
                                      FT rho, FT N_lcl) {
    const FT eM = eps_2M<FT>(), eN = eps_2M<FT>();
    const FT kcc = acnv.kcc, x_star = acnv.x_star, rho0 = acnv.rho0, A = acnv.A, a = acnv.a, b = acnv.b;
    const FT nu_c = pdf_c.nu_c;
    FT safe_q_lcl = jmax(q_lcl, eM);
    FT safe_N_lcl = jmax(N_lcl, eN);
    FT L_lcl = rho * safe_q_lcl;
    FT x_lcl = jmin(x_star, L_lcl / safe_N_lcl);
    FT safe_q_rai = jmax(FT(0), q_rai);
    FT tau = 1 - safe_q_lcl / (safe_q_lcl + safe_q_rai);  // Eq. (5)
    FT phi_au = (q_rai < eM) ? FT(0) : A * pow_(tau, a) * pow_(1 - pow_(tau, a), b);
    FT nu1 = nu_c + 1;
    FT dL_rai_dt = kcc / 20 / x_star * (nu_c + 2) * (nu_c + 4) / (nu1 * nu1) * (L_lcl * L_lcl) *
                   (x_lcl * x_lcl) * (1 + phi_au / ((1 - tau) * (1 - tau))) * rho0 / rho;  // Eq. (4)
    FT dN_rai_dt = dL_rai_dt / x_star;
    FT dL_lcl_dt = -dL_rai_dt;
    FT dN_lcl_dt = -2 * dN_rai_dt;
    bool cond = (q_lcl < eM) || (N_lcl < eN);
    LclRaiRates<FT> r;
    r.dq_lcl_dt = cond ? FT(0) : dL_lcl_dt / rho;
    r.dN_lcl_dt = cond ? FT(0) : dN_lcl_dt;
    r.dq_rai_dt = cond ? FT(0) : dL_rai_dt / rho;
    r.dN_rai_dt = cond ? FT(0) : dN_rai_dt;
    return r;
}

// CM2.accretion(::SB2006, ...)                                     CM2:445-470
template <class FT>
inline LclRaiRates<FT> accretion(const typename PT<FT>::sb_accr& accr, FT q_lcl, FT q_rai, FT rho,
                                 FT N_lcl) {
    const FT eM = eps_2M<FT>(), eN = eps_2M<FT>();
    FT safe_q_lcl = jmax(q_lcl, eM);
    FT safe_q_rai = jmax(q_rai, eM);
    FT safe_N_lcl = jmax(N_lcl, eN);
    FT L_lcl = rho * safe_q_lcl;
    FT L_rai = rho * safe_q_rai;
    FT x_lcl = L_lcl / safe_N_lcl;
    FT tau = 1 - safe_q_lcl / (safe_q_lcl + safe_q_rai);  // Eq. (5)
    FT phi_ac = pow_(tau / (tau + accr.tau0), accr.c);  // Eq. (8)
    FT dL_rai_dt = accr.kcr * L_lcl * L_rai * phi_ac * sqrt_(accr.rho0 / rho);  // Eq. (7)
    FT dL_lcl_dt = -dL_rai_dt;
    FT dN_lcl_dt = dL_lcl_dt / x_lcl;
    bool cond = (q_lcl < eM) || (q_rai < eM) || (N_lcl < eN);
    LclRaiRates<FT> r;
    r.dq_lcl_dt = cond ? FT(0) : dL_lcl_dt / rho;
    r.dN_lcl_dt = cond ? FT(0) : dN_lcl_dt;
    r.dq_rai_dt = cond ? FT(0) : dL_rai_dt / rho;
    r.dN_rai_dt = FT(0);
    return r;
}

// CM2.cloud_liquid_self_collection                                 CM2:488-501
template <class FT>
inline FT cloud_liquid_self_collection(const typename PT<FT>::sb_acnv& acnv,
                                       const typename PT<FT>::sb_pdf_c& pdf_c, FT q_lcl, FT rho,
                                       FT dN_lcl_dt_au) {
    FT L_lcl = rho * q_lcl;
    FT v = -acnv.kcc * (pdf_c.nu_c + 2) / (pdf_c.nu_c + 1) * (acnv.rho0 / rho) * (L_lcl * L_lcl) -
           dN_lcl_dt_au;
    return (q_lcl < eps_2M<FT>()) ? FT(0) : v;
}

// CM2.rain_self_collection                                         CM2:545-560
template <class FT>
inline FT rain_self_collection(const typename PT<FT>::sb_pdf_r& pdf, const typename PT<FT>::sb_self& self,
                               FT q_rai, FT rho, FT N_rai) {
    const FT eM = eps_2M<FT>(), eN = eps_2M<FT>();
    FT safe_q = jmax(q_rai, eM);
    FT safe_N = jmax(N_rai, eN);
    FT L_rai = rho * safe_q;
    FT Ar, Br;
    pdf_rain_parameters_mass<FT>(pdf, safe_q, rho, safe_N, Ar, Br);
    FT v = -self.krr * N_rai * L_rai * sqrt_(pdf.rho0 / rho) * pow_(1 + self.kappa_rr / Br, self.d);
    bool cond = (q_rai < eM) || (N_rai < eN);
    return cond ? FT(0) : v;
}

// CM2.rain_breakup                                                 CM2:579-601
template <class FT>
inline FT rain_breakup(const typename PT<FT>::sb_pdf_r& pdf, const typename PT<FT>::sb_brek& brek,
                       FT q_rai, FT rho, FT N_rai, FT dN_rai_dt_sc) {
    const FT eM = eps_2M<FT>(), eN = eps_2M<FT>();
    FT safe_q = jmax(q_rai, eM);
    FT safe_N = jmax(N_rai, eN);
    RainPDF<FT> r = pdf_rain_parameters<FT>(pdf, safe_q, rho, safe_N);
    FT Dr = cbrt_(r.xr_mean * 6 / (pi<FT>() * pdf.rho_w));
    FT dD = Dr - brek.Deq;
    FT phi = (Dr < brek.Dr_th) ? FT(-1)
                               : ((Dr <= brek.Deq) ? brek.kbr * dD : exp_(brek.kappa_br * dD) - 1);
    FT v = -(phi + 1) * dN_rai_dt_sc;  // Eq. (13)
    bool cond = (q_rai < eM) || (N_rai < eN);
    return cond ? FT(0) : v;
}

// CM2.Γ_incl                                                      CM2:746-753
template <class FT> inline FT Gamma_incl(FT a, FT x) {
    return exp_(-x) / ((FT(0.33) - FT(0.7) * a) * pow_(x, FT(0.08) - FT(0.93) * a) +
                           (FT(1.34) - FT(0.1) * a) * pow_(x, FT(0.8) - a));
}

// CM2.rain_evaporation                                             CM2:780-828
template <class FT>
inline void rain_evaporation(const typename PT<FT>::sb2006& sb, const typename PT<FT>::air& aps,
                             const Thermo<FT>& tps, FT q_tot, FT q_lcl, FT q_icl, FT q_rai, FT q_sno,
                             FT rho, FT N_rai, FT T, FT& dNrho_dt, FT& dq_dt) {
    const FT eM = eps_2M<FT>(), eN = eps_2M<FT>();
    FT S = tps.supersat_liq(q_tot, q_lcl + q_rai, q_icl + q_sno, rho, T);
    const auto& evap = sb.evap;
    FT rho_w = sb.pdf_r.rho_w;
    FT x_star = sb.pdf_r.xr_min;
    FT G = G_func_liquid<FT>(aps, tps, T);
    FT safe_q = jmax(q_rai, eM);
    FT safe_N = jmax(N_rai, eN);
    RainPDF<FT> r = pdf_rain_parameters<FT>(sb.pdf_r, safe_q, rho, safe_N);
    FT xr_mean = r.xr_mean;
    FT Dr = cbrt_(6 * xr_mean / (pi<FT>() * rho_w));
    FT t_star = cbrt_(FT(6) * x_star / xr_mean);
    FT a_vent_0 = evap.a_vent_0_coeff * Gamma_incl<FT>(FT(-1), t_star);
    FT b_vent_0 = evap.b_vent_0_coeff * Gamma_incl<FT>(evap.beta_vent_0, t_star);
    FT a_vent_1 = evap.a_vent_1;
    FT b_vent_1 = evap.b_vent_1;
    FT N_Re = evap.alpha * pow_(xr_mean, evap.beta) * sqrt_(evap.rho0 / rho) * Dr / aps.nu_air;
    FT cbrt_Sc = cbrt_(aps.nu_air / jmax(aps.D_vapor, eps_numerics<FT>()));
    FT sqrt_N_Re = sqrt_(N_Re);
    FT Fv0 = a_vent_0 + b_vent_0 * cbrt_Sc * sqrt_N_Re;
    FT Fv1 = a_vent_1 + b_vent_1 * cbrt_Sc * sqrt_N_Re;
    FT dn = jmin(FT(0), 2 * pi<FT>() * G * S * N_rai * Dr * Fv0 / xr_mean);
    FT dq = jmin(FT(0), 2 * pi<FT>() * G * S * N_rai * Dr * Fv1 / rho);
    dNrho_dt = ((q_rai < eM) || (xr_mean / x_star < eps<FT>()) || (N_rai <= eN) || (S >= 0)) ? FT(0) : dn;
    dq_dt = ((q_rai < eM) || (N_rai <= eN) || (S >= 0)) ? FT(0) : dq;
}

// CM2.number_tendency_from_mass_limits                             CM2:882-891
template <class FT> inline FT number_tendency_from_mass_limits(FT x_min, FT x_max, FT tau, FT q, FT n) {
    FT n_target = (q < eps_2M<FT>()) ? FT(0) : jclamp(n, q / x_max, q / x_min);
    return (n_target - n) / tau;
}

// CM2.cloud_terminal_velocity                                      CM2:647-664
template <class FT>
inline void cloud_terminal_velocity(const typename PT<FT>::sb_pdf_c& pdf_c,
                                    const typename PT<FT>::vel_stokes& vel, FT q_liq, FT rho, FT N_liq,
                                    FT& vt0, FT& vt1) {
    const FT eM = eps_2M<FT>(), eN = eps_2M<FT>();
    FT safe_q = jmax(q_liq, eM);
    FT safe_N = jmax(N_liq, eN);
    FT Ac, Bc;
    pdf_cloud_parameters_mass<FT>(pdf_c, safe_q, rho, safe_N, Ac, Bc);
    FT t = FT(6) / vel.rho_w / pi<FT>();
    FT pref = FT(1.0 / 18) * cbrt_(t * t) * (vel.rho_w / rho - 1) * vel.grav / vel.nu_air;
    FT v0 = pref * generalized_gamma_Mn<FT>(pdf_c.nu_c, pdf_c.mu_c, Bc, safe_N, FT(2.0 / 3)) / safe_N;
    FT v1 = pref * generalized_gamma_Mn<FT>(pdf_c.nu_c, pdf_c.mu_c, Bc, safe_N, FT(5.0 / 3)) / rho / safe_q;
    bool cond = (N_liq < eN) || (q_liq < eM);
    vt0 = cond ? FT(0) : v0;
    vt1 = cond ? FT(0) : v1;
}

// CM2.rain_terminal_velocity(::SB2006, ::SB2006VelType, ...)       CM2:685-702, helpers 720-739
template <class FT>
inline void rain_terminal_velocity_sb(const typename PT<FT>::sb_pdf_r& pdf_r,
                                      const typename PT<FT>::vel_sb2006& vel, FT q_rai, FT rho, FT N_rai,
                                      FT& vt0, FT& vt1) {
    const FT eM = eps_2M<FT>(), eN = eps_2M<FT>();
    FT safe_q = jmax(q_rai, eM);
    FT safe_N = jmax(N_rai, eN);
    RainPDF<FT> r = pdf_rain_parameters<FT>(pdf_r, safe_q, rho, safe_N);
    FT Dr_mean = r.Dr_mean;
    FT pa0, pb0, pa1, pb1;
    if (pdf_r.limited) {
        pa0 = pb0 = pa1 = pb1 = FT(1);
    } else {
        FT lam = 1 / Dr_mean;
        FT rc = -1 / (2 * vel.cR) * log_(vel.aR / vel.bR);
        auto G1 = [](FT t) { return exp_(-t); };
        auto G4 = [](FT t) { return (t * t * t + 3 * (t * t) + 6 * t + 6) * exp_(-t); };
        pa0 = G1(2 * rc * lam);
        pb0 = G1(2 * rc * (lam + vel.cR));
        pa1 = G4(2 * rc * lam) / 6;
        pb1 = G4(2 * rc * (lam + vel.cR)) / 6;
    }
    FT s = sqrt_(vel.rho0 / rho);
    FT d1 = 1 + vel.cR * Dr_mean;
    FT d2 = d1 * d1;
    FT v0 = jmax(FT(0), s * (vel.aR * pa0 - vel.bR * pb0 / d1));
    FT v1 = jmax(FT(0), s * (vel.aR * pa1 - vel.bR * pb1 / (d2 * d2)));
    vt0 = (N_rai < eN) ? FT(0) : v0;
    vt1 = (q_rai < eM) ? FT(0) : v1;
}

// CM2.rain_terminal_velocity(::SB2006, ::Chen2022VelTypeRain, ...) CM2:703-719
template <class FT>
inline void rain_terminal_velocity_chen(const typename PT<FT>::sb_pdf_r& pdf_r,
                                        const typename PT<FT>::vel_chen_rain& vel, FT q_rai, FT rho,
                                        FT N_rai, FT& vt0, FT& vt1) {
    const FT eM = eps_2M<FT>(), eN = eps_2M<FT>();
    FT aiu[3], bi[3], ciu[3];
    chen2022_vel_coeffs_rain<FT>(vel, rho, aiu, bi, ciu);
    FT safe_q = jmax(q_rai, eM);
    FT safe_N = jmax(N_rai, eN);
    RainPDF<FT> r = pdf_rain_parameters<FT>(pdf_r, safe_q, rho, safe_N);
    FT v0 = 0, v3 = 0;
    for (int i = 0; i < 3; ++i) v0 += chen2022_exponential_pdf<FT>(aiu[i], bi[i], ciu[i], r.Dr_mean, 0);
    for (int i = 0; i < 3; ++i) v3 += chen2022_exponential_pdf<FT>(aiu[i], bi[i], ciu[i], r.Dr_mean, 3);
    vt0 = (N_rai < eN) ? FT(0) : jmax(FT(0), v0);
    vt1 = (q_rai < eM) ? FT(0) : jmax(FT(0), v3);
}

// ---- BulkMicrophysicsTendencies.jl: warm_rain_tendencies_2m (BMT:707-782) ------
template <class FT> struct Warm2MOut {
    FT dq_lcl_dt, dn_lcl_dt, dq_rai_dt, dn_rai_dt;
    FT leaf[CUMICRO_SB2006_NLEAF];
};

template <class FT>
inline Warm2MOut<FT> warm_rain_tendencies_2m(const typename PT<FT>::params_2m_warm& p, FT T, FT q_tot,
                                             FT q_lcl, FT q_rai, FT q_ice, FT rho, FT n_lcl, FT n_rai) {
    Thermo<FT> tps(p.tps);
    const auto& sb = p.sb;
    Warm2MOut<FT> o;
    FT N_lcl = rho * n_lcl;
    FT N_rai = rho * n_rai;
    FT dq_lcl_dt = 0, dq_rai_dt = 0, dn_lcl_dt = 0, dn_rai_dt = 0;

    // condensation / evaporation of cloud liquid                    BMT:730-739
    FT cond = conv_q_vap_to_q_lcl_const<FT>(p.condevap_tau_relax, tps, q_tot, q_lcl, q_ice, q_rai,
                                            FT(0), rho, T);
    dq_lcl_dt += cond;
    dn_lcl_dt += FT(0);
    // rain evaporation                                             BMT:741-744
    FT ev_n, ev_q;
    rain_evaporation<FT>(sb, p.aps, tps, q_tot, q_lcl, q_ice, q_rai, FT(0), rho, N_rai, T, ev_n, ev_q);
    dq_rai_dt += ev_q;
    dn_rai_dt += ev_n / rho;
    // autoconversion                                               BMT:746-751
    LclRaiRates<FT> ac = autoconversion<FT>(sb.acnv, sb.pdf_c, q_lcl, q_rai, rho, N_lcl);
    dq_lcl_dt += ac.dq_lcl_dt;
    dq_rai_dt += ac.dq_rai_dt;
    dn_lcl_dt += ac.dN_lcl_dt / rho;
    dn_rai_dt += ac.dN_rai_dt / rho;
    // cloud liquid self-collection                                 BMT:753-755
    FT sc = cloud_liquid_self_collection<FT>(sb.acnv, sb.pdf_c, q_lcl, rho, ac.dN_lcl_dt);
    dn_lcl_dt += sc / rho;
    // accretion                                                    BMT:757-761
    LclRaiRates<FT> accr = accretion<FT>(sb.accr, q_lcl, q_rai, rho, N_lcl);
    dq_lcl_dt += accr.dq_lcl_dt;
    dq_rai_dt += accr.dq_rai_dt;
    dn_lcl_dt += accr.dN_lcl_dt / rho;
    // rain self-collection, breakup                                BMT:763-769
    FT rsc = rain_self_collection<FT>(sb.pdf_r, sb.self, q_rai, rho, N_rai);
    dn_rai_dt += rsc / rho;
    FT rbr = rain_breakup<FT>(sb.pdf_r, sb.brek, q_rai, rho, N_rai, rsc);
    dn_rai_dt += rbr / rho;
    // number adjustment                                            BMT:771-779
    FT adj_l = number_tendency_from_mass_limits<FT>(sb.pdf_c.xc_min, sb.pdf_c.xc_max, sb.numadj_tau, q_lcl, n_lcl);
    dn_lcl_dt += adj_l;
    FT adj_r = number_tendency_from_mass_limits<FT>(sb.pdf_r.xr_min, sb.pdf_r.xr_max, sb.numadj_tau, q_rai, n_rai);
    dn_rai_dt += adj_r;

    o.dq_lcl_dt = dq_lcl_dt;
    o.dn_lcl_dt = dn_lcl_dt;
    o.dq_rai_dt = dq_rai_dt;
    o.dn_rai_dt = dn_rai_dt;
    o.leaf[CUMICRO_SB_COND_DQ_LCL] = cond;
    o.leaf[CUMICRO_SB_EVAP_DN_RAI] = ev_n;
    o.leaf[CUMICRO_SB_EVAP_DQ_RAI] = ev_q;
    o.leaf[CUMICRO_SB_ACNV_DQ_LCL] = ac.dq_lcl_dt;
    o.leaf[CUMICRO_SB_ACNV_DN_LCL] = ac.dN_lcl_dt;
    o.leaf[CUMICRO_SB_ACNV_DQ_RAI] = ac.dq_rai_dt;
    o.leaf[CUMICRO_SB_ACNV_DN_RAI] = ac.dN_rai_dt;
    o.leaf[CUMICRO_SB_LCL_SELFCOL] = sc;
    o.leaf[CUMICRO_SB_ACCR_DQ_LCL] = accr.dq_lcl_dt;
    o.leaf[CUMICRO_SB_ACCR_DN_LCL] = accr.dN_lcl_dt;
    o.leaf[CUMICRO_SB_ACCR_DQ_RAI] = accr.dq_rai_dt;
    o.leaf[CUMICRO_SB_RAI_SELFCOL] = rsc;
    o.leaf[CUMICRO_SB_RAI_BREAKUP] = rbr;
    o.leaf[CUMICRO_SB_NUMADJ_LCL] = adj_l;
    o.leaf[CUMICRO_SB_NUMADJ_RAI] = adj_r;
    return o;
}

// BMT:820-854 bulk_microphysics_tendencies(::Microphysics2Moment, mp{WR,Nothing}, ...)
template <class FT>
inline Warm2MOut<FT> bmt2m_warm(const typename PT<FT>::params_2m_warm& p, FT rho, FT T, FT q_tot,
                                FT q_lcl, FT n_lcl, FT q_rai, FT n_rai) {
    rho = clamp_to_nonneg(rho);
    q_tot = clamp_to_nonneg(q_tot);
    q_lcl = clamp_to_nonneg(q_lcl);
    q_rai = clamp_to_nonneg(q_rai);
    n_lcl = clamp_to_nonneg(n_lcl);
    n_rai = clamp_to_nonneg(n_rai);
    return warm_rain_tendencies_2m<FT>(p, T, q_tot, q_lcl, q_rai, FT(0), rho, n_lcl, n_rai);
}

// ---- alternative closures: CM2.conv_q_lcl_to_q_rai / accretion for KK2000, B1994, TC1980, LD2004     CM2:920-1002
template <class FT, class A> inline FT alt_2m(const A& p, int what, bool smooth, FT q_lcl, FT q_rai, FT rho, FT N_d) {
    switch (what) {
        case 0: { q_lcl = jmax(FT(0), q_lcl); return FT(p.kk_acnv_A) * pow_(q_lcl, FT(p.kk_acnv_a)) * pow_(N_d, FT(p.kk_acnv_b)) * pow_(rho, FT(p.kk_acnv_c)); }
        case 1: {
            q_lcl = jmax(FT(0), q_lcl);
            FT d;
            if (smooth) {
                FT lo = logistic_function<FT>(N_d, FT(p.b_acnv_N_0), FT(p.b_acnv_k));
                FT hi = 1 - lo;
                d = lo * FT(p.b_acnv_d_low) + hi * FT(p.b_acnv_d_high);
            } else d = (N_d >= FT(p.b_acnv_N_0)) ? FT(p.b_acnv_d_low) : FT(p.b_acnv_d_high);
            return FT(p.b_acnv_C) * pow_(d, FT(p.b_acnv_a)) * pow_(FT(q_lcl * rho), FT(p.b_acnv_b)) * pow_(N_d, FT(p.b_acnv_c)) / rho;
        }
        case 2: {
            q_lcl = jmax(FT(0), q_lcl);
            FT thr = FT(p.tc_acnv_m0_liq_coeff) * N_d / rho * pow_(FT(p.tc_acnv_r_0), FT(p.tc_acnv_me_liq));
            FT o = smooth ? logistic_function<FT>(q_lcl, thr, FT(p.tc_acnv_k)) : FT((q_lcl - thr > FT(0)) ? 1 : 0);
            return FT(p.tc_acnv_D) * pow_(q_lcl, FT(p.tc_acnv_a)) * pow_(N_d, FT(p.tc_acnv_b)) * o;
        }
        case 3: {
            if (q_lcl <= eps_2M<FT>()) return FT(0);
            FT r_vol = cbrt_(3 * q_lcl * rho / 4 / pi<FT>() / FT(p.ld_rho_w) / N_d) * 1000000;
            FT b6 = cbrt_((r_vol + 3) / r_vol);
            FT b2 = b6 * b6;
            FT E = FT(p.ld_E_0) * (b2 * b2 * b2);
            FT R6 = b6 * r_vol;
            FT R6C = FT(p.ld_R_6C_0) / cbrt_(sqrt_(FT(q_lcl * rho))) / sqrt_(R6);
            FT o = smooth ? logistic_function<FT>(R6, R6C, FT(p.ld_k)) : FT((R6 - R6C > FT(0)) ? 1 : 0);
            FT L = q_lcl * rho;
            return E * (L * L * L) / N_d / rho * o;
        }
        case 4: { q_lcl = jmax(FT(0), q_lcl); q_rai = jmax(FT(0), q_rai); return FT(p.kk_accr_A) * pow_(FT(q_lcl * q_rai), FT(p.kk_accr_a)) * pow_(rho, FT(p.kk_accr_b)); }
        case 5: { q_lcl = jmax(FT(0), q_lcl); q_rai = jmax(FT(0), q_rai); return FT(p.b_accr_A) * q_lcl * rho * q_rai; }
        case 6: { q_lcl = jmax(FT(0), q_lcl); q_rai = jmax(FT(0), q_rai); return FT(p.tc_accr_A) * q_lcl * q_rai; }
        default: return FT(0);
    }
}

}  // namespace orc
