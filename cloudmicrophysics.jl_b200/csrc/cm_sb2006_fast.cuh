// cm_sb2006_fast.cuh — the headline kernel body: BMT.bulk_microphysics_tendencies(::Microphysics2Moment, mp{WR,Nothing}, ...)
// (BMT:707-782, 820-854) for the reference's default SB2006 block STRUCTURE (exponents acnv.b = 3, accr.c = 4, self.d = -5;
// every value is still a run-time parameter), four tendencies out, written for the B200 FP64 pipe:
// the figure of merit is the number of issued instructions (FP64 and not), DESIGN.md §3.1.
//
// What differs from the general form in cm_sb2006.cuh (which stays the leaf / generic path and the parity cross-check):
//   * limited rain PSD (CM2:87-110) in LOG SPACE: Eq. (94)-(97) are clamps of affine combinations of log L and log N, so the two
//     logarithms replace three reciprocals, a cube root, a fourth root and the logarithm of xr_mean; only xr_mean is consumed here
//     (lam, N0r are outputs of the terminal-velocity entry points, not of the tendencies);
//   * every power of xr_mean is exp_(c log xr_mean + c'): parameter-only factors (2 pi, a_vent, b_vent cbrt(Sc), sqrt(alpha/nu) ...)
//     are folded into the additive constants and denominators on the host (make_w2k);
//   * log_abs_ (8 FP64) where a logarithm is an additive term; the two-piece exp_ (9 FP64);
//   * thresholds against eps(FT) = 2^-52 / 2^-23 are integer compares of the high word (the operands are clamped to >= +0);
//   * tau = 1 - q/(q + q_r) through the correctly rounded shared-reciprocal quotient (bit-identical to the IEEE division, no
//     slow-path subroutine), one reciprocal for x_lcl and 1/x_lcl, one for the two Γ_incl denominators.
// Regime predicates are the reference's (same operators on the same quantities).  __host__ __device__: tests/native runs the same
// code on the CPU against the oracle.
#pragma once
#include "cm_sb2006.cuh"

namespace cm {

// Host-derived constants of one launch (doubles; the functor computes in Float64 for both method types).
struct W2K {   // (fields in the order the body reads them: neighbours share one 128-bit constant-bank load)
    int eps_hi, _pad;         // high word of eps (low word is zero: eps is a power of two)
    double eps, eps_n;        // eps(FT), cbrt(floatmin(FT)) of the METHOD's float type
    // thermodynamics (cm_thermo.cuh)
    double inv_T_triple, T_triple, a_liq, b_liq, press_triple, dcp_vl, Lv0, R_v, inv_R_v;   // Lv = dcp_vl T + Lv0
    double dcp_lv, dcp_vd, cp_d, dcp_iv, tau_cond;
    // rain PSD: logs of the limiter bounds; log(pi rho_w)/3, /4; 1/3, -1/3 (full precision)
    double lxmin, lxmax, nthird, c3, lN0min, lN0max, c4, llmin, llmax, third;
    double cDr, rho0;         // Dr = cDr xr^(1/3); sqrt(rho0/rho)
    double inv_K, RvD;        // 1/max(K_therm, eps_n), R_v/max(D_vapor, eps_n)
    // ventilation table (limited PSD; see W2Tab): u = tab_inv_h lx + tab_u0 is the interval coordinate of log xr_mean
    double tab_inv_h, tab_u0;
    double kvx, cvx, av1_Dr;  // 2 pi b_vent_1 cbrt(Sc) Dr sqrt(N_Re) / r4 = exp(kvx lx + cvx), kvx = kv + 1/3; av1 cDr
    // autoconversion / accretion / self-collection / breakup
    double x_star, acnv_a, acnv_A, acnv_pref, inv_x_star, lclsc_pref;
    double tau0, kcr;         // kcr sqrt(accr.rho0 / pdf_r.rho0)
    double krc, nkrr;         // kappa_rr cbrt(1/6), -krr
    double Deq, kappa_br, Dr_th, kbr;
    double inv_xc_max, inv_xc_min, inv_xr_max, inv_xr_min, inv_tau_adj;
    // closed-form evaporation (no table / not-limited PSD)
    double lt0, ct;           // log t* = lt0 - lx/3,  t* = ct xr^(-1/3)
    double ne1[2], de[2], c1[2], c2[2];   // Γ_incl(a_k, t) = exp(-t + ne1 log t) / (c1 + c2 t^de), c1, c2 pre-divided (see make_w2k)
    double kv, cv;            // sqrt(N_Re) = exp(kv lx + cv) (rho0/rho)^(1/4)
    double av1, bv1;          // 2 pi a_vent_1, 2 pi b_vent_1 cbrt(Sc)
    double pi_rho_w;
};


// ---- ventilation table ------------------------------------------------------------------------------------------------------
// Under the limited PSD log xr_mean lives in [log xr_min, log xr_max], and the number-tendency side of CM2.rain_evaporation depends
// on it through two smooth functions only (t = cbrt(6 x*/xr), everything else is parameters):
//     T1(lx) = Dr/xr 2 pi a_vent_0_coeff Γ_incl(-1, t)            T2(lx) = Dr/xr 2 pi b_vent_0_coeff cbrt(Sc) Γ_incl(beta_vent_0, t) sqrt(N_Re) / r4
//     dN/dt = G S N_rai (T1 + T2 r4),   r4 = (rho0/rho)^(1/4)
// i.e. four real powers of t, exp(-t) and a division per point (4 exp_, 2 reciprocals: ~60 FP64 instructions).  The host tabulates
// T1 and T2 for the parameter block at hand: kTabN intervals of log xr_mean, degree-7 polynomials in the centred interval
// coordinate s in [-1/2, 1/2] (Chebyshev interpolation in long double, converted to monomials), verified against the closed form
// before use (max relative error < 1e-15, else the table is not used and the closed form runs).  The functions are analytic in
// log xr (nearest singularity ~12 away from the real axis against an interval half-width of 0.08), so degree 7 converges to ~2e-16.
// Per point: 3 + 14 FP64 instructions and 8 128-bit shared-memory loads.
constexpr int kTabN = 64;            // intervals; rows = kTabN + 1 (nodes at the interval centres i h, i = 0..kTabN)
constexpr int kTabRow = 18;          // doubles per row: 8 + 8 coefficients + 2 of padding (144 B: rows rotate over the banks)
constexpr int kTabDoubles = (kTabN + 1) * kTabRow;

struct W2TabEval {   // closed forms in long double (host only)
    long double xr_min, cDr, two_pi_a0, two_pi_b0_sc, c1[2], c2[2], e1[2], e2[2], alpha_nu, beta, rho_ratio4;
    void operator()(long double lx, long double& T1, long double& T2) const {
        const long double xr = expl(lx), cx = expl(lx / 3), Dr = cDr * cx;
        const long double t = cbrtl(6 * xr_min / xr);
        long double g[2];
        for (int i = 0; i < 2; ++i) g[i] = expl(-t) / (c1[i] * powl(t, e1[i]) + c2[i] * powl(t, e2[i]));
        const long double sqrt_NRe = sqrtl(alpha_nu * powl(xr, beta) * Dr) * rho_ratio4;   // without (pdf_r.rho0/rho)^(1/4)
        T1 = Dr / xr * two_pi_a0 * g[0];
        T2 = Dr / xr * two_pi_b0_sc * g[1] * sqrt_NRe;
    }
};

// Fills tab[kTabDoubles] and k.tab_*; returns the verified max relative error (the caller uses the table only if it is < 1e-15).
inline double build_w2_table(const cumicro_params_2m_warm_f64& p, W2K& k, double* tab) {
    const auto& sb = p.sb;
    const long double pi = 3.141592653589793238462643383279502884L;
    W2TabEval f;
    f.xr_min = sb.pdf_r.xr_min;
    f.cDr = cbrtl(6 / (pi * (long double)sb.pdf_r.rho_w));
    const long double D = std::max((long double)p.aps.D_vapor, (long double)k.eps_n);
    f.two_pi_a0 = 2 * pi * sb.evap.a_vent_0_coeff;
    f.two_pi_b0_sc = 2 * pi * sb.evap.b_vent_0_coeff * cbrtl((long double)p.aps.nu_air / D);
    const long double a[2] = {-1.0L, (long double)sb.evap.beta_vent_0};
    for (int i = 0; i < 2; ++i) {   // the literals of CM2:746-753 as the Float64 method reads them
        f.c1[i] = (long double)0.33 - (long double)0.7 * a[i];
        f.c2[i] = (long double)1.34 - (long double)0.1 * a[i];
        f.e1[i] = (long double)0.08 - (long double)0.93 * a[i];
        f.e2[i] = (long double)0.8 - a[i];
    }
    f.alpha_nu = (long double)sb.evap.alpha / p.aps.nu_air;
    f.beta = sb.evap.beta;
    f.rho_ratio4 = powl((long double)sb.evap.rho0 / sb.pdf_r.rho0, 0.25L);
    const long double lo = logl((long double)sb.pdf_r.xr_min), hi = logl((long double)sb.pdf_r.xr_max);
    const long double h = (hi - lo) / kTabN;
    k.tab_inv_h = (double)(1 / h);
    k.tab_u0 = (double)(-lo / h);
    // Chebyshev nodes on [-1/2, 1/2], interpolation in the Chebyshev basis (discrete orthogonality), then T_j(2 s) -> monomials in s
    constexpr int M = 8;
    long double xs[M], Tm[M][M];   // Tm[j][q]: coefficient of s^q in T_j(2 s)
    for (int q = 0; q < M; ++q) xs[q] = 0.5L * cosl(pi * (2 * q + 1) / (2 * M));
    for (int j = 0; j < M; ++j) for (int q = 0; q < M; ++q) Tm[j][q] = 0;
    Tm[0][0] = 1; Tm[1][1] = 2;
    for (int j = 2; j < M; ++j)
        for (int q = 0; q < M; ++q) Tm[j][q] = (q > 0 ? 4 * Tm[j - 1][q - 1] : 0) - Tm[j - 2][q];   // T_j(y) = 2 y T_{j-1} - T_{j-2}, y = 2 s
    for (int i = 0; i <= kTabN; ++i) {
        long double v[2][M];
        for (int q = 0; q < M; ++q) f(lo + (i + xs[q]) * h, v[0][q], v[1][q]);
        for (int fn = 0; fn < 2; ++fn) {
            long double cheb[M], mono[M];
            for (int j = 0; j < M; ++j) {
                long double acc = 0;
                for (int q = 0; q < M; ++q) acc += v[fn][q] * cosl(pi * j * (2 * q + 1) / (2 * M));
                cheb[j] = acc * (j == 0 ? 1.0L : 2.0L) / M;
            }
            for (int q = 0; q < M; ++q) { mono[q] = 0; for (int j = 0; j < M; ++j) mono[q] += cheb[j] * Tm[j][q]; }
            for (int q = 0; q < M; ++q) tab[i * kTabRow + fn * M + q] = (double)mono[q];
        }
        tab[i * kTabRow + 16] = tab[i * kTabRow + 17] = 0.0;
    }
    // verification: Float64 Horner against the closed form at points that are not interpolation nodes
    double worst = 0.0;
    for (int i = 0; i <= kTabN; ++i)
        for (int m = 0; m <= 8; ++m) {
            const double sv = -0.5 + m / 8.0;
            if ((i == 0 && sv < 0) || (i == kTabN && sv > 0)) continue;
            long double t1, t2;
            f(lo + (i + (long double)sv) * h, t1, t2);
            for (int fn = 0; fn < 2; ++fn) {
                const double* c = tab + i * kTabRow + fn * M;
                double acc = c[M - 1];
                for (int q = M - 2; q >= 0; --q) acc = fma(acc, sv, c[q]);
                const long double tr = fn ? t2 : t1;
                worst = std::max(worst, (double)fabsl(((long double)acc - tr) / tr));
            }
        }
    return worst;
}

// STD structure (see warm2m_fast): exponents 3 / 4 / -5 and strictly positive ventilation coefficients.
inline bool w2k_supported(const cumicro_params_2m_warm_f64& p) {
    const auto& sb = p.sb;
    auto pos = [](double x) { return x > 0.0 && x < 1.7e308; };
    return sb.acnv.b == 3.0 && sb.accr.c == 4.0 && sb.self.d == -5.0 && pos(sb.evap.a_vent_0_coeff) && pos(sb.evap.b_vent_0_coeff) &&
           pos(sb.evap.alpha) && pos(p.aps.nu_air) && pos(sb.pdf_r.rho_w) && pos(sb.pdf_r.xr_min) && pos(sb.pdf_r.xr_max) &&
           pos(sb.pdf_r.N0_min) && pos(sb.pdf_r.N0_max) && pos(sb.pdf_r.lam_min) && pos(sb.pdf_r.lam_max) && pos(sb.pdf_r.rho0) &&
           pos(sb.evap.rho0) && pos(sb.accr.rho0) && pos(p.aps.D_vapor) && pos(p.aps.K_therm) && pos(sb.evap.b_vent_1) &&
           sb.pdf_r.xr_min < sb.pdf_r.xr_max;
}

inline W2K make_w2k(const cumicro_params_2m_warm_f64& p, bool method_is_f32) {
    W2K k{};
    const auto& t = p.tps;
    const auto& sb = p.sb;
    const double pi = 3.141592653589793238462643383279502884;
    k.eps = method_is_f32 ? 1.1920928955078125e-07 : 2.220446049250313e-16;
    k.eps_n = method_is_f32 ? 2.2737367544323206e-13 : 2.8126442852362996e-103;
    k.eps_hi = method_is_f32 ? 0x3E800000 : 0x3CB00000;
    k.inv_T_triple = 1.0 / t.T_triple; k.T_triple = t.T_triple;
    const double dcp_vl = t.cp_v - t.cp_l;
    k.a_liq = dcp_vl / t.R_v; k.b_liq = (t.LH_v0 - dcp_vl * t.T_0) / t.R_v;
    k.press_triple = t.press_triple; k.R_v = t.R_v; k.inv_R_v = 1.0 / t.R_v;
    k.dcp_vl = dcp_vl; k.Lv0 = t.LH_v0 - dcp_vl * t.T_0;
    k.cp_d = t.cp_d; k.dcp_vd = t.cp_v - t.cp_d; k.dcp_lv = t.cp_l - t.cp_v; k.dcp_iv = t.cp_i - t.cp_v;
    k.tau_cond = p.condevap_tau_relax;
    k.inv_K = 1.0 / std::max(p.aps.K_therm, k.eps_n);
    k.RvD = t.R_v / std::max(p.aps.D_vapor, k.eps_n);
    const double prw = pi * sb.pdf_r.rho_w;
    k.lxmin = std::log(sb.pdf_r.xr_min); k.lxmax = std::log(sb.pdf_r.xr_max);
    k.lN0min = std::log(sb.pdf_r.N0_min); k.lN0max = std::log(sb.pdf_r.N0_max);
    k.llmin = std::log(sb.pdf_r.lam_min); k.llmax = std::log(sb.pdf_r.lam_max);
    k.c3 = std::log(prw) / 3; k.c4 = std::log(prw) / 4;
    k.third = 1.0 / 3.0; k.nthird = -1.0 / 3.0;
    k.pi_rho_w = prw; k.rho0 = sb.pdf_r.rho0;
    const double six_x_star = 6.0 * sb.pdf_r.xr_min;
    k.lt0 = std::log(six_x_star) / 3; k.ct = std::cbrt(six_x_star);
    const double cbrt_Sc = std::cbrt(p.aps.nu_air / std::max(p.aps.D_vapor, k.eps_n));
    k.cDr = std::cbrt(6.0 / prw);
    // sqrt(N_Re) = sqrt(alpha xr^beta sqrt(evap.rho0/rho) Dr / nu) = exp(kv lx + cv) (pdf_r.rho0/rho)^(1/4)
    k.kv = 0.5 * sb.evap.beta + 1.0 / 6.0;
    k.cv = 0.5 * std::log(sb.evap.alpha * k.cDr / p.aps.nu_air) + 0.25 * std::log(sb.evap.rho0 / sb.pdf_r.rho0);
    // 2 pi Fv0 = 2 pi a_vent_0_coeff Γ_incl(-1, t) + 2 pi b_vent_0_coeff cbrt(Sc) Γ_incl(beta_vent_0, t) sqrt(N_Re): the prefactors divide c1, c2
    const double a[2] = {-1.0, sb.evap.beta_vent_0};
    const double pref[2] = {2 * pi * sb.evap.a_vent_0_coeff, 2 * pi * sb.evap.b_vent_0_coeff * cbrt_Sc};
    for (int i = 0; i < 2; ++i) {
        const double e1 = 0.08 - 0.93 * a[i];
        k.ne1[i] = -e1;
        k.de[i] = (0.8 - a[i]) - e1;
        k.c1[i] = (0.33 - 0.7 * a[i]) / pref[i];
        k.c2[i] = (1.34 - 0.1 * a[i]) / pref[i];
    }
    k.av1 = 2 * pi * sb.evap.a_vent_1; k.bv1 = 2 * pi * sb.evap.b_vent_1 * cbrt_Sc;
    k.inv_xr_min = 1.0 / sb.pdf_r.xr_min; k.inv_xr_max = 1.0 / sb.pdf_r.xr_max;
    k.inv_xc_min = 1.0 / sb.pdf_c.xc_min; k.inv_xc_max = 1.0 / sb.pdf_c.xc_max;
    k.inv_tau_adj = 1.0 / sb.numadj_tau;
    const double nu_c = sb.pdf_c.nu_c;
    k.x_star = sb.acnv.x_star; k.inv_x_star = 1.0 / sb.acnv.x_star;
    k.acnv_pref = sb.acnv.kcc / 20 / sb.acnv.x_star * (nu_c + 2) * (nu_c + 4) / ((nu_c + 1) * (nu_c + 1)) * sb.acnv.rho0;
    k.acnv_a = sb.acnv.a; k.acnv_A = sb.acnv.A;
    k.lclsc_pref = sb.acnv.kcc * (nu_c + 2) / (nu_c + 1) * sb.acnv.rho0;
    k.tau0 = sb.accr.tau0; k.kcr = sb.accr.kcr * std::sqrt(sb.accr.rho0 / sb.pdf_r.rho0);
    k.krc = sb.self.kappa_rr * std::cbrt(1.0 / 6.0); k.nkrr = -sb.self.krr;
    k.Deq = sb.brek.Deq; k.Dr_th = sb.brek.Dr_th; k.kbr = sb.brek.kbr; k.kappa_br = sb.brek.kappa_br;
    k.tab_inv_h = 0.0; k.tab_u0 = 0.0;
    k.av1_Dr = k.av1 * k.cDr;
    k.kvx = k.kv + 1.0 / 3.0;
    k.cvx = k.cv + std::log(k.bv1 * k.cDr);
    return k;
}

// x < eps(FT) for x >= +0 (or NaN: false), eps a power of two
CM_HD bool lt_eps_(double x, int eps_hi) { return hi32(x) < eps_hi; }
CM_HD double neg_(double x) { return mk64(hi32(x) ^ (int)0x80000000, lo32(x)); }

// LIM = 1 / 0: limited / not-limited rain PSD.  x = (rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai), q_ice as seen by the thermodynamics.
// TAB: `tab` (shared memory on the device) holds the verified ventilation table of this parameter block (limited PSD only).
template <int LIM, bool TAB = false>
CM_HD void warm2m_fast(const W2K& k, double rho, double T, double q_tot, double q_lcl, double n_lcl, double q_rai, double n_rai,
                       double q_ice, bool have_ice, double (&y)[4], const double* tab = nullptr) {
    static_assert(!(TAB && LIM == 0), "the ventilation table covers the bounded log xr_mean of the limited PSD");
    const int eh = k.eps_hi;
    // input clamps                                                   BMT:827-836
    rho = clamp0_(rho); q_tot = clamp0_(q_tot); q_lcl = clamp0_(q_lcl);
    q_rai = clamp0_(q_rai); n_lcl = clamp0_(n_lcl); n_rai = clamp0_(n_rai);
    const double N_lcl = rho * n_lcl, N_rai = rho * n_rai;   // BMT:718-719
    const double inv_rho = rcp_(rho);
    const bool ql_off = lt_eps_(q_lcl, eh), qr_off = lt_eps_(q_rai, eh), Nl_off = lt_eps_(N_lcl, eh), Nr_off = lt_eps_(N_rai, eh);

    // ---- thermodynamic state (TDI:60-125)
    const double inv_T = rcp_(T);
    const double log_Tr = log_abs_(T * k.inv_T_triple);
    const double dinvT = (T - k.T_triple) * inv_T * k.inv_T_triple;
    const double p_vs = k.press_triple * exp_(fma(k.a_liq, log_Tr, k.b_liq * dinvT));
    const double inv_p_vs = rcp_(fmax_(p_vs, k.eps_n));
    const double Lv = fma(k.dcp_vl, T, k.Lv0);
    const double q_liq = q_lcl + q_rai;
    const double qv = have_ice ? clamp0_(q_tot - q_liq - q_ice) : clamp0_(q_tot - q_liq);
    const double rho_Rv_T = rho * k.R_v * T;
    const double qv_sat = p_vs * rcp_(rho_Rv_T);
    const double sat_excess = qv - qv_sat;
    const double LT = Lv * inv_T;
    const double g1 = fma(LT, k.inv_R_v, -1.0);

    // ---- NEQ._conv_q_vap_to_q_lcl_const                            NEQ:117-140
    double cond;
    {
        double cp_air = fma(k.dcp_lv, q_liq, fma(k.dcp_vd, q_tot, k.cp_d));
        if (have_ice) cp_air = fma(k.dcp_iv, q_ice, cp_air);
        const double dqsl_dT = (qv_sat * inv_T) * g1;                                    // NEQ.dqcld_dT
        const double inv_ts = cp_air * rcp_(k.tau_cond * fma(Lv, dqsl_dT, cp_air));    // 1/(tau Gamma)
        const double nq = neg_(q_lcl);
        const double m = (sat_excess < nq) ? nq : sat_excess;   // se < 0 ? -min(-se, q_lcl) : se
        cond = m * inv_ts;
    }

    // ---- rain size distribution: log xr_mean, xr_mean^(±1/3)          CM2:67-110
    const double sq_rai = qr_off ? k.eps : q_rai;
    const double sN_rai = Nr_off ? k.eps : N_rai;
    const double L_rai = rho * sq_rai;
    double lx, cx, inv_cx = 0.0, xr_ratio = 1.0;
    if (LIM == 1) {
        const double lL = log_abs_(L_rai), lN = log_abs_(sN_rai);
        const double lxt = clamp_(lL - lN, k.lxmin, k.lxmax);                              // Eq. (94)
        const double lN0 = clamp_(fma(lxt, k.nthird, lN + k.c3), k.lN0min, k.lN0max);      // Eq. (95)
        const double llam = clamp_(fma(lN0 - lL, 0.25, k.c4), k.llmin, k.llmax);           // Eq. (96)
        lx = clamp_((lL - lN0) + llam, k.lxmin, k.lxmax);                                  // Eq. (97)
        cx = exp_(lx * k.third);
        if (!TAB) inv_cx = rcp_(cx);
    } else {
        const double xr_mean = L_rai * rcp_(sN_rai);
        xr_ratio = xr_mean * k.inv_xr_min;
        cx = cbrt_pair_(xr_mean, inv_cx);
        lx = log_abs_(xr_mean);
    }
    const double Dr = cx * k.cDr;                               // CM2:590, 802
    const double sqrt_rho0_rho = sqrtp_(k.rho0 * inv_rho);

    // ---- CM2.rain_evaporation                                       CM2:780-828
    double evap_dn, evap_dq;
    {
        const double S = fma(qv * rho_Rv_T, inv_p_vs, -1.0);                    // TDI.supersaturation_over_liquid
        const double G = rcp_(fma(LT * k.inv_K, g1, (T * inv_p_vs) * k.RvD));   // CO.G_func_liquid
        const double r4 = sqrtp_(sqrt_rho0_rho);                                // (rho0/rho)^(1/4)
        // gates: q_rai < eps || N_rai <= eps zero the common factor (every other factor is finite: safe values); S >= 0 makes it
        // non-negative and min(0, .) returns the reference's 0                                                   CM2:822-827
        const bool off_q = qr_off || (N_rai <= k.eps);
        const double common = off_q ? 0.0 : G * S * N_rai;
        double dn;
        if (TAB) {
            // dN/dt = G S N_rai (T1 + T2 r4) from the ventilation table;  dq/dt = G S N_rai / rho (2 pi a_vent_1 Dr + 2 pi b_vent_1 cbrt(Sc) Dr sqrt(N_Re))
            const double magic = 6755399441055744.0;
            const double u = fma(lx, k.tab_inv_h, k.tab_u0);
            const double tm = u + magic;
            const double sv = u - (tm - magic);                    // in [-1/2, 1/2]
            const double* row = tab + lo32(tm) * kTabRow;
            double t1 = row[7], t2 = row[15];
#pragma unroll
            for (int q = 6; q >= 0; --q) { t1 = fma(t1, sv, row[q]); t2 = fma(t2, sv, row[8 + q]); }
            dn = cap0_(common * fma(t2, r4, t1));
            const double w1 = exp_(fma(k.kvx, lx, k.cvx)) * r4;     // 2 pi b_vent_1 cbrt(Sc) Dr sqrt(N_Re)
            evap_dq = cap0_(common * inv_rho * fma(k.av1_Dr, cx, w1));
        } else {
            const double inv_xr_mean = inv_cx * inv_cx * inv_cx;
            const double lt = fma(lx, k.nthird, k.lt0);
            const double t_star = k.ct * inv_cx;
            const double a0 = fma(k.ne1[0], lt, -t_star), a1 = fma(k.ne1[1], lt, -t_star);
            const double E0 = (LIM == 1) ? exp_(a0) : exp_full_(a0);
            const double E1 = (LIM == 1) ? exp_(a1) : exp_full_(a1);
            const double den0 = fma(k.c2[0], exp_(k.de[0] * lt), k.c1[0]);
            const double den1 = fma(k.c2[1], exp_(k.de[1] * lt), k.c1[1]);
            const double vv = exp_(fma(k.kv, lx, k.cv)) * r4;                           // sqrt(N_Re)
            const double Fv0 = fma(E1 * den0, vv, E0 * den1) * rcp_(den0 * den1);       // 2 pi (a_vent_0 + b_vent_0 cbrt(Sc) sqrt(N_Re))
            const double Fv1 = fma(k.bv1, vv, k.av1);
            const double cD = common * Dr;
            dn = cap0_(cD * Fv0 * inv_xr_mean);
            evap_dq = cap0_(cD * Fv1 * inv_rho);
        }
        evap_dn = (LIM == 0 && xr_ratio < k.eps) ? 0.0 : dn;
    }

    // ---- CM2.autoconversion, cloud_liquid_self_collection, accretion   CM2:396-501
    double dq_au, dNr_au, dNl_sum, dq_ac;
    const double Lsq = L_rai * sqrt_rho0_rho;
    {
        const double sq_lcl = ql_off ? k.eps : q_lcl;
        const double sN_lcl = Nl_off ? k.eps : N_lcl;
        const double L_lcl = rho * sq_lcl;
        const double LL = L_lcl * L_lcl;
        const double xl = L_lcl * rcp_(sN_lcl);     // L_lcl / N_lcl
        const double xlc = (xl < k.x_star) ? xl : k.x_star;
        const double s = sq_lcl + q_rai;
        const double tau = 1.0 - divr_(sq_lcl, s, rcp_cr_(s));      // SB2006 Eq. (5), the IEEE quotient
        const double omt = 1.0 - tau;
        const double tau_a = exp_(k.acnv_a * log_abs_(tau));
        const double oma = 1.0 - tau_a;
        double phi = k.acnv_A * tau_a * (oma * oma * oma);
        phi = qr_off ? 0.0 : phi;
        const bool off = ql_off || Nl_off;
        // one gate on dL: its multiples are exact zeros too (inv_rho, 1/x* finite)
        const double dL = off ? 0.0 : k.acnv_pref * LL * (xlc * xlc) * fma(phi, rcp_(omt * omt), 1.0) * inv_rho;   // Eq. (4)
        dNr_au = dL * k.inv_x_star;
        dq_au = dL * inv_rho;
        const double dNl_au = -2.0 * dNr_au;
        const double sc_l = ql_off ? 0.0 : (-(k.lclsc_pref * inv_rho) * LL - dNl_au);    // CM2:488-501 (q_lcl >= eps: rho q_lcl = L_lcl)
        const double pa = tau * rcp_(tau + k.tau0);
        const double pa2 = pa * pa;
        // dL_rai = kcr L_lcl L_rai phi sqrt(rho0/rho) (Eq. 7, 8);  dN_lcl = -dL_rai / x_lcl = -(kcr L_rai phi sqrt(rho0/rho)) N_lcl   CM2:462-466
        const double acc = (off || qr_off) ? 0.0 : k.kcr * Lsq * (pa2 * pa2);
        dq_ac = acc * L_lcl * inv_rho;
        const double dNl_ac = -acc * sN_lcl;
        dNl_sum = (dNl_au + sc_l) + dNl_ac;
    }

    // ---- CM2.rain_self_collection / rain_breakup                       CM2:545-601
    double sc, br;
    {
        const double a = fma(k.krc, cx, 1.0);
        const double a2 = a * a;
        const double v = k.nkrr * N_rai * Lsq * rcp_(a2 * a2 * a);
        const bool no_rain = qr_off || Nr_off;
        sc = no_rain ? 0.0 : v;
        const double dD = Dr - k.Deq;
        const double ex = (LIM == 1) ? exp_(k.kappa_br * dD) : exp_full_(k.kappa_br * dD);
        const double phi_p1 = (Dr < k.Dr_th) ? 0.0 : ((Dr <= k.Deq) ? fma(k.kbr, dD, 1.0) : ex);
        br = -phi_p1 * sc;                                             // Eq. (13): -(Φ_br + 1) dN_sc
        if (LIM == 0) br = no_rain ? 0.0 : br;
    }

    // ---- number adjustment (Horn 2012)                                 BMT:771-779, CM2:882-891
    const double tl = ql_off ? 0.0 : clamp_(n_lcl, q_lcl * k.inv_xc_max, q_lcl * k.inv_xc_min);
    const double tr = qr_off ? 0.0 : clamp_(n_rai, q_rai * k.inv_xr_max, q_rai * k.inv_xr_min);
    const double na_l = (tl - n_lcl) * k.inv_tau_adj;
    const double na_r = (tr - n_rai) * k.inv_tau_adj;

    // ---- aggregate in the order of BMT:736-779
    y[0] = (cond - dq_au) - dq_ac;
    y[1] = fma(dNl_sum, inv_rho, na_l);
    y[2] = (evap_dq + dq_au) + dq_ac;
    y[3] = fma(((evap_dn + dNr_au) + sc) + br, inv_rho, na_r);
}

}  // namespace cm
