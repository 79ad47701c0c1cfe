// cm_p3.cuh — P3 ice scheme: particle properties, size distribution, incomplete-gamma
// utilities, quadrature-based process rates.
//
// Device form of src/P3_particle_properties.jl (:43-475), P3_size_distribution.jl (:8-237),
// P3_integral_properties.jl (:34-45), P3_terminal_velocity.jl (:4-173), P3_processes.jl
// (:64-712), src/Quadrature.jl (:62-125), UT.gamma_inc / gamma_inc_inv (UT:92-252), the
// regularised ratios (UT:445-509) and the PSD closures / bounds of CM2 (CM2:176-355).
//
// Work decomposition (one WARP per ice-bearing grid point).  Every integral of
// the reference is a fixed-order quadrature over up to 4 mass-regime segments, so one point is
// ~2e6 FP64 instructions of perfectly regular work:
//   * the outer quadrature nodes (4 segments x n) are spread over the 32 lanes, partial sums
//     are combined with warp shuffles;
//   * everything that does not depend on the outer ice diameter is evaluated ONCE per point and
//     staged in shared memory: the cloud / rain inner nodes (D, w n(D), m(D), v(D)), and for the
//     closed-form rain integral the 24 (z, α) incomplete-gamma pairs at the two fixed ends of
//     the rain spectrum together with Γ(z)/α^z (the reference re-evaluates 96 gamma_inc per outer
//     node; 24 remain, at the crossover diameter);
//   * all powers D^b are exp(b log D) from one logarithm per node.
// The fixed iteration counts of the reference (30 / 20 gamma_inc terms, 10 / 8 Brent steps) are
// kept: their truncation error is part of the reference's result.
#pragma once
#include "cm_sb2006.cuh"

namespace cm {

constexpr int kQuadMax = 128;
constexpr int kGam = 24;          // (z, α) pairs of closed_rain_inner_NM: 4 velocity terms x (p + i) in 0..5

// ---- compact special functions for the per-point set-up -------------------------------------------
// The set-up of one point (thresholds, PSD normalisation, Γ tables) runs once per ~80 k warp-instructions, but the
// CUDA libm expansions of pow / lgamma / tgamma / log it would inline are ~8000 SASS instructions (130 KB): walking
// them evicts the hot quadrature loops from the 32 KB instruction cache on every point.  These versions are a few
// hundred instructions in total, out of line, and accurate to ~1e-15 on the argument ranges of the scheme.
#include "cm_gamma_poly.inc"
static const double cm_gamma_poly_host[CM_GAMMA_POLY_DEG + 1] = CM_GAMMA_POLY_INIT;
#ifdef __CUDACC__
static __constant__ double cm_gamma_poly_dev[CM_GAMMA_POLY_DEG + 1] = CM_GAMMA_POLY_INIT;
#endif
// Γ(x) for x >= 1: Γ(x) = (x-1)(x-2)...(x-k) Γ(x-k), x-k in [1, 2), 1/Γ on [1, 2] by a degree-15 polynomial.
__host__ __device__ __noinline__ inline double tgamma_pos_(double x) {
    if (!(x >= 1.0) || x > 64.0) return tgamma(x);
    double p = 1.0;
    while (x >= 2.0) { x -= 1.0; p *= x; }
    const double u = 2.0 * x - 3.0;
#ifdef __CUDA_ARCH__
    const double* c = cm_gamma_poly_dev;
#else
    const double* c = cm_gamma_poly_host;
#endif
    double q = c[CM_GAMMA_POLY_DEG];
#pragma unroll
    for (int i = CM_GAMMA_POLY_DEG - 1; i >= 0; --i) q = fma(q, u, c[i]);
    return p / q;
}
__host__ __device__ __noinline__ inline double lgamma_pos_(double x) {
    if (!(x >= 1.0) || x > 64.0) return lgamma(x);
    return log(tgamma_pos_(x));
}
__host__ __device__ __noinline__ inline double exp_nl_(double x) { return exp_full_(x); }            // one copy of exp for cold callers
__host__ __device__ __noinline__ inline double expm1_nl_(double x) { return expm1(x); }
__host__ __device__ __noinline__ inline double log1p_nl_(double x) { return log1p(x); }
__host__ __device__ __noinline__ inline double logp_nl_(double x) {     // one copy of logp_ for cold callers
    return (x > 2.3e-308 && x < 1.7e308) ? logp_(x) : log(x);
}

// x^y, x positive normal: integer part of y by multiplication, fractional part through exp/log, so that the
// relative error stays ~1e-15 for the large exponents of Γ(z)/α^z (|y ln x| up to ~100)
__host__ __device__ __noinline__ inline double pow_pos_(double x, double y) {
    if (!(x > 2.3e-308 && x < 1.7e308)) return pow(x, y);   // 0, Inf, NaN, subnormal (e.g. ρ_g = 0 -> D_gr = Inf): libm semantics
    const double yi = floor(y);
    const int n = (fabs(yi) <= 64.0) ? (int)yi : 0;
    const double f = y - (double)n;
    double r = 1.0, b = (n < 0) ? 1.0 / x : x;
    for (int i = (n < 0 ? -n : n); i > 0; i >>= 1) { if (i & 1) r *= b; b *= b; }
    return (f == 0.0) ? r : r * exp_nl_(f * logp_nl_(x));
}
// ---- UT.gamma_inc: series (x < a + 1) or continued fraction, the reference's fixed iteration counts   UT:92-144
// Both branches are division-free restatements of the reference's loops (one division at the end instead of one / two per
// step): they sum the same terms / reach the same convergent, so they agree with it to rounding (<= 6e-15 relative, measured
// against a Float64 transcription of UT's loops over a in [0.5, 12], x to 1e4) including where 30 steps have not converged.
// Unrolling further is a loss (2^19 points: CF x1 47.35 ms, CF x2 51.3): the kernel stays code-size bound.
#ifndef P3_CF_UNROLL
#define P3_CF_UNROLL 1
#endif
#ifndef P3_SERIES_TEST_EVERY
#define P3_SERIES_TEST_EVERY 4   /* exit test of the series every 2nd / 4th term: process rates 49.98 / 48.95 ms per 2^20 points */
#endif
constexpr int kP3CfUnroll = P3_CF_UNROLL;
struct PQ { double P, Q; };
// The closing quotient of a gamma_inc evaluation.  The divisors (the scaled series product P_K >= a, the scaled continued-fraction
// numerator A_K) are normal and finite, so the Markstein quotient through a correctly rounded reciprocal (cm_math.cuh: the IEEE
// result, except for divisors with an all-ones significand) replaces CUDA's division and its slow-path branch.
#ifndef P3_DIV_MARKSTEIN
#define P3_DIV_MARKSTEIN 1   /* process rates, 2^20 points: 48.95 ms with CUDA's division, 48.29 ms with the Markstein quotient */
#endif
CM_HD double p3_div_(double x, double d) {
#if P3_DIV_MARKSTEIN
    return divr_(x, d, rcp_cr_(d));
#else
    return div_(x, d);
#endif
}
// gamma_inc_core_: the series / continued fraction with the prefactor x^a e^-x / Γ(a) given.  A caller that evaluates the orders
// a, a + 1, a + 2, ... at one x (the closed-form rain integrals: six consecutive orders per velocity term) forms the first
// prefactor with one exponential (from a logarithm of x it already has: gamma_inc_factor_) and the next ones by
// factor(a + 1) = factor(a) x / a — Γ(a + 1) = a Γ(a) — instead of five more exponentials; rounding-level differences.
__host__ __device__ __noinline__ inline PQ gamma_inc_core_(double a, double x, double factor, int iters) {
    PQ r;
    if (x <= 0.0) { r.P = 0.0; r.Q = 1.0; return r; }
    if (x == num<double>::inf()) { r.P = 1.0; r.Q = 0.0; return r; }
    if (x < a + 1.0) {
        // Σ_k x^k / (a (a+1) ... (a+k)) = S_K / P_K with S_k = S_{k-1} (a+k) + x^k, P_k = P_{k-1} (a+k): the reference's term
        // recurrence (term *= x / (a+k); sum += term) without its division per term — same sum, rounding-level difference.
        // Magnitudes stay below (a+K)^K: a < 1e5 at K = 30 is safe; P3 calls this with a <= μ_max + 1 and the Chen exponents + 7.
        // Exit test term < sum * 5.5e-17 on the scaled pair, every P3_SERIES_TEST_EVERY-th term (a converged term adds less than half an ulp).
        double P = a, X = 1.0, S = 1.0, ak = a;
        int k = 1;
#if P3_SERIES_TEST_EVERY == 4
#pragma unroll 1
        for (; k + 3 <= iters; k += 4) {   // exit test every fourth term: a warp runs to its slowest lane anyway, and a converged term adds nothing
            ak += 1.0; P *= ak; X *= x; S = fma(S, ak, X);
            ak += 1.0; P *= ak; X *= x; S = fma(S, ak, X);
            ak += 1.0; P *= ak; X *= x; S = fma(S, ak, X);
            ak += 1.0; P *= ak; X *= x; S = fma(S, ak, X);
            if (X < S * 5.5e-17) { k = iters + 1; break; }
        }
#pragma unroll 1
        for (; k <= iters; ++k) { ak += 1.0; P *= ak; X *= x; S = fma(S, ak, X); }
#else
#pragma unroll 1
        for (; k + 1 <= iters; k += 2) {
            ak += 1.0; P *= ak; X *= x; S = fma(S, ak, X);
            ak += 1.0; P *= ak; X *= x; S = fma(S, ak, X);
            if (X < S * 5.5e-17) { k = iters + 1; break; }
        }
        if (k <= iters) { ak += 1.0; P *= ak; X *= x; S = fma(S, ak, X); }
#endif
        r.P = clamp_(factor * p3_div_(S, P), 0.0, 1.0);
        r.Q = 1.0 - r.P;
    } else {
        // Legendre's continued fraction 1 / (b_0 + a_1 / (b_1 + a_2 / (b_2 + ...))), a_k = -k (k - a), b_k = x + 2k + 1 - a, cut
        // after `iters` partial quotients like the reference's modified-Lentz loop (UT:155-180), evaluated by the three-term
        // recurrence A_k = b_k A_{k-1} + a_k A_{k-2} of its numerator and denominator: no division per step (Lentz: two), the
        // same convergent in exact arithmetic (its 1e-30 guards never engage for x >= a + 1), <= 6e-15 relative apart in
        // Float64.  Both sequences carry s^k, s = the power of two below 1 / b_iters, so they stay O(1) for any x.
        const double b0 = x + 1.0 - a;
        const double top = b0 + 2.0 * (double)iters;
        const double sc = mk64((2045 - ((hi32(top) >> 20) & 0x7ff)) << 20, 0);
        const double sc2 = sc + sc;
        double A2 = 1.0, A1 = b0 * sc, B2 = 0.0, B1 = sc;
        double bs = b0 * sc, ks = 0.0, kas = -a * sc;
#pragma unroll kP3CfUnroll
        for (int k = 1; k <= iters; ++k) {
            bs += sc2; ks += sc; kas += sc;
            const double at = ks * kas;                 // -a_k s^2
            const double A = fma(bs, A1, -(at * A2));
            const double B = fma(bs, B1, -(at * B2));
            A2 = A1; A1 = A; B2 = B1; B1 = B;
        }
        r.Q = clamp_(factor * p3_div_(B1, A1), 0.0, 1.0);
        r.P = 1.0 - r.Q;
    }
    return r;
}
__host__ __device__ inline double gamma_inc_factor_(double a, double x, double lx, double lga) { return exp_nl_(a * lx - x - lga); }
__host__ __device__ inline PQ gamma_inc_(double a, double x, double lga, int iters) {
    const bool ok = x > 0.0 && x < num<double>::inf();
    return gamma_inc_core_(a, x, ok ? gamma_inc_factor_(a, x, logp_nl_(x), lga) : 0.0, iters);
}

// ---- UT.gamma_inc_inv: Halley, <= 15 steps with the reference's exits                UT:205-252
__host__ __device__ __noinline__ inline double gamma_inc_inv_(double a, double p, double q, int iters, double eps) {
    if (p <= 0.0) return 0.0;
    if (q <= 0.0) return num<double>::inf();
    double x = (p < 0.5) ? pow_pos_(p * tgamma_pos_(a + 1.0), 1.0 / a) : (a - logp_nl_(q));
    const bool use_q = p > 0.5;
    const double lga = lgamma_pos_(a);
#pragma unroll 1
    for (int i = 1; i <= 15; ++i) {
        const PQ g = gamma_inc_(a, x, lga, iters);
        const double f = use_q ? g.Q - q : g.P - p;
        double fprime = exp_nl_((a - 1.0) * logp_nl_(x) - x - lga);
        fprime = use_q ? -fprime : fprime;
        if (fprime == 0.0) break;
        const double f2 = (a - 1.0 - x) / x;
        double step = f / (fprime * (1.0 - 0.5 * f * f2 / fprime));
        if (x - step <= 0.0) step = 0.5 * x;
        x = x - step;
        if (fabs(step) < eps * x) break;
    }
    return x;
}

// ---- UT.sgs_weight_function / _regularised_ratio                                   UT:445-485
CM_HD double regularised_ratio_(double num_, double den, double eps) {
    double w;
    if (den < 0.0) w = 0.0;
    else if (den > fmin(1.0, 42.0 * eps)) w = 1.0;
    else if (4.0 * den < eps) w = 0.0;
    else w = (1.0 + tanh(2.0 * atanh(1.0 - 2.0 * pow(1.0 - den, -1.0 / log2(1.0 - eps))))) / 2.0;
    return (den < eps * eps) ? 0.0 : w * num_ / den;
}

// ---- host-derived constants of one launch ------------------------------------------------------
struct P3K {
    // ParametersP3
    double alpha_va, beta_va, gamma_a, sigma_a;
    double slope_a, slope_b, slope_c, mu_max, mu_const;
    double vent_a, vent_b, rim_a, rim_b, rim_c, rim_rho_ice, rim8;
    double tau_wet, rho_i, rho_l08, T_freeze;
    int slope_power_law, aspect_oblate;
    // thresholds: (6 α_va / (π ρ))^(1/(3 - β_va))
    double thr_p, thr_coef, D_th;
    double pi6, phi_coef;                 // π/6, 3 sqrt(π)
    double log_pi4, log_gamma_a, half_log_pi, lphc_i;   // log(π/4), log γ, log(π)/2, log(3 sqrt(π) / (4 ρᵢ))
    double log_rho_i_pi6, log_alpha_va, log_mu_c, log3;
    // air
    double cbrt_Nsc, inv_nu_air, K_therm, D_vapor;
    // Chen 2022 ice tables at ρᵢ = 916.7 (hard-coded in the reference, P3_terminal_velocity.jl:32)
    double As, Bs, Cs, Es, Fs, Gs1000, Al, Bl, Cl, El, Fl, Gl1000, Hl, cutoff, pow1000_Cl, pow1000_Fl;
    // liquid PSDs
    double rho_w, mliq_coef, m_shd, log_km;
    double nu_cD, mu_cD, cloud_log_z_lo, cloud_log_z_hi;      // log gamma_inc_inv((νcD+1)/μcD, p | 1-p), p = 1e-5
    double rain_cll_lo, rain_cll_hi;                        // cloglog(p), cloglog(1 - p)
    // Bigg / F23
    double het_a, het_B, tau_act, m_nuc, V1, cloud_M3_ratio, cloud_M6_ratio, frost_b, frost_log_a, frost_T_freeze;
    double subdep_tau;
    int n, gamma_iters, brent_iters;
    double eps;
};

__host__ inline P3K make_p3_k(const cumicro_params_p3_f64& p, bool method_is_f32) {
    P3K k{};
    const double pi = 3.141592653589793;
    const auto& s = p.scheme;
    k.eps = method_is_f32 ? 1.1920928955078125e-07 : 2.220446049250313e-16;
    k.gamma_iters = method_is_f32 ? 20 : 30;
    k.brent_iters = method_is_f32 ? 8 : 10;
    k.alpha_va = s.alpha_va; k.beta_va = s.beta_va; k.gamma_a = s.gamma; k.sigma_a = s.sigma;
    k.slope_a = s.slope_a; k.slope_b = s.slope_b; k.slope_c = s.slope_c; k.mu_max = s.slope_mu_max; k.mu_const = s.slope_mu_const;
    k.vent_a = s.vent_a; k.vent_b = s.vent_b; k.rim_a = s.rim_a; k.rim_b = s.rim_b; k.rim_c = s.rim_c; k.rim_rho_ice = s.rim_rho_ice;
    k.rim8 = s.rim_a + s.rim_b * 8.0 + s.rim_c * (8.0 * 8.0);
    k.tau_wet = s.tau_wet; k.rho_i = s.rho_i; k.rho_l08 = 0.8 * s.rho_l; k.T_freeze = s.T_freeze;
    k.slope_power_law = s.slope_power_law; k.aspect_oblate = s.aspect_oblate;
    k.thr_p = 1.0 / (3.0 - s.beta_va);
    k.thr_coef = 6.0 * s.alpha_va;
    k.D_th = std::pow(6.0 * s.alpha_va / (pi * s.rho_i), 1.0 / (3.0 - s.beta_va));
    k.pi6 = pi / 6.0;
    k.log_rho_i_pi6 = std::log(s.rho_i * pi / 6.0); k.log_alpha_va = std::log(s.alpha_va);
    k.log_mu_c = std::log(p.warm.sb.pdf_c.mu_c); k.log3 = std::log(3.0);
    k.phi_coef = 3.0 * std::sqrt(pi);
    k.log_pi4 = std::log(pi / 4.0); k.log_gamma_a = std::log(s.gamma); k.half_log_pi = 0.5 * std::log(pi);
    k.lphc_i = std::log(k.phi_coef / (4.0 * s.rho_i));
    const auto& aps = p.warm.aps;
    k.cbrt_Nsc = std::cbrt(aps.nu_air / aps.D_vapor);
    k.inv_nu_air = 1.0 / aps.nu_air; k.K_therm = aps.K_therm; k.D_vapor = aps.D_vapor;
    {   // CO.Chen2022_vel_coeffs small / large ice, ρᵢ-only parts                      CO:302-349
        const double ri = 916.7, l = std::log(ri), sq = std::sqrt(ri);
        const auto& c = p.vel_small_ice;
        k.As = c.A[1] * (l * l) - c.A[2] * l + c.A[0];
        k.Bs = 1.0 / (c.B[0] + c.B[1] * l + c.B[2] / sq);
        k.Cs = c.C[0] + c.C[1] * std::exp(c.C[2] * ri) + c.C[3] * sq;
        k.Es = c.E[0] - c.E[1] * (l * l) + c.E[2] * sq;
        k.Fs = -std::exp(c.F[0] - c.F[1] * (l * l) + c.F[2] * l);
        k.Gs1000 = 1.0 / (c.G[0] + c.G[1] / l - c.G[2] * l / ri) * 1000.0;
        const auto& g = p.vel_large_ice;
        k.Al = g.A[0] + g.A[1] * l + g.A[2] / (ri * sq);
        k.Bl = std::exp(g.B[0] + g.B[1] * (l * l) + g.B[2] * l);
        k.Cl = std::exp(g.C[0] + g.C[1] / l + g.C[2] / ri);
        k.El = g.E[0] + g.E[1] * l * sq + g.E[2] * sq;
        k.Fl = g.F[0] + g.F[1] * l - std::exp(std::log(-g.F[2]) - ri);
        k.Gl1000 = 1.0 / (g.G[0] + g.G[1] * l * sq + g.G[2] / sq) * 1000.0;
        k.Hl = g.H[0] + g.H[1] * (ri * ri) * sq + std::exp(std::log(-g.H[2]) - ri);
        k.cutoff = c.cutoff;
        k.pow1000_Cl = std::exp(k.Cl * 6.907755278982137); k.pow1000_Fl = std::exp(k.Fl * 6.907755278982137);
    }
    const auto& pc = p.warm.sb.pdf_c;
    k.rho_w = pc.rho_w;
    k.mliq_coef = pc.rho_w;                                   // m_liq(D) = ρw (D³ π / 6)
    k.m_shd = pc.rho_w * (1e-3 * 1e-3 * 1e-3 * pi / 6.0);
    k.log_km = std::log(pc.rho_w * pi / 6.0);
    k.nu_cD = 3.0 * pc.nu_c + 2.0;
    k.mu_cD = 3.0 * pc.mu_c;
    {
        const double pq = 0.00001, Y1 = 1.0 - pq, a = (k.nu_cD + 1.0) / k.mu_cD;
        k.cloud_log_z_lo = std::log(gamma_inc_inv_(a, pq, 1.0 - pq, k.gamma_iters, k.eps));
        k.cloud_log_z_hi = std::log(gamma_inc_inv_(a, Y1, 1.0 - Y1, k.gamma_iters, k.eps));
        k.rain_cll_lo = std::log(-std::log1p(-pq));
        k.rain_cll_hi = std::log(-std::log1p(-Y1));
        k.cloud_M3_ratio = std::tgamma((k.nu_cD + 1.0 + 3.0) / k.mu_cD) / std::tgamma((k.nu_cD + 1.0) / k.mu_cD);
        k.cloud_M6_ratio = std::tgamma((k.nu_cD + 1.0 + 6.0) / k.mu_cD) / std::tgamma((k.nu_cD + 1.0) / k.mu_cD);
    }
    k.het_a = p.rain_freezing_het_a; k.het_B = p.rain_freezing_het_B; k.tau_act = p.tau_act;
    k.m_nuc = s.rho_i * (10e-6 * 10e-6 * 10e-6 * pi / 6.0);
    k.V1 = pi / 6.0;
    k.frost_b = p.ice_nucleation.b; k.frost_log_a = p.ice_nucleation.log_a; k.frost_T_freeze = p.ice_nucleation.T_freeze;
    k.subdep_tau = p.warm.subdep_tau_relax;
    k.n = p.quad.n;
    return k;
}

#ifdef __CUDACC__
CM_DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
CM_DEV double bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Per-warp scratch in shared memory (sized for the launch's quadrature order n).
struct P3Scratch {
    double* cD;    // cloud inner nodes: diameter
    double* cWN;   //   w_j n_c(D_j)
    double* cM;    //   m_liq(D_j)
    double* cV;    //   v_liq(D_j)
    double* rD;    // rain inner nodes
    double* rWNM;  //   w_j n_r(D_j) m_liq(D_j)
    double* rV;
    double* gz;    // [kGam] z_k
    double* glg;   // [kGam] loggamma(z_k)
    double* gG;    // [kGam] Γ(z_k) / α_k^z_k
    double* gP0;   // [kGam] P, Q at α_k D_min and α_k D_max
    double* gQ0;
    double* gP1;
    double* gQ1;
    double* tA;    // [4] α_j = λ_r + c_j of the 4 velocity terms (j = 0: the v_i term)
    double* tC;    // [4] a_j (j >= 1)
    __host__ __device__ static constexpr int doubles(int n) { return 7 * n + 7 * kGam + 8; }
    __device__ void bind(double* base, int n) {
        cD = base; cWN = cD + n; cM = cWN + n; cV = cM + n; rD = cV + n; rWNM = rD + n; rV = rWNM + n;
        gz = rV + n; glg = gz + kGam; gG = glg + kGam; gP0 = gG + kGam; gQ0 = gP0 + kGam; gP1 = gQ0 + kGam; gQ1 = gP1 + kGam; tA = gQ1 + kGam; tC = tA + 4;
    }
};

// One ice-bearing grid point, evaluated cooperatively by the 32 lanes of a warp.  All members are
// warp-uniform.
struct P3Point {
    // inputs
    double rho, T, L_ice, N_ice, F_rim, rho_rim;
    // P3State thresholds                                        P3_particle_properties.jl:43-56
    double rho_g, D_gr, D_cr, lphc_g;   // lphc_g = log(3 sqrt(π) / (4 ρ_g))
    // mass regime coefficients: a D^b as exp(la + b log D)
    double la_small, la_unr, la_grp, la_part;
    // PSD
    double mu, logN0, lam;
    // Chen velocity curves at this air density
    double sa0, sa1, sb, sc1;          // small ice: a0 D^b + a1 D^b e^{-c1 D}
    double ga0, ga1, gb0, gb1, gc1;    // large ice
    double ra[3], rb[3], rc[3];        // rain / cloud liquid

    CM_DEV int regime(double D, const P3K& k) const {
        return (D < k.D_th) ? 0 : ((F_rim == 0.0) ? 1 : ((D < D_gr) ? 2 : ((D < D_cr) ? 3 : 4)));
    }
    // everything the integrands need at one ice diameter
    // One code path for all regimes and both velocity curves (coefficients are selected, not branched on): the body
    // stays within the ~6 KB L0 instruction cache of an SM sub-partition, which is what bounds this kernel.
    struct Node { double mass, dmass_dD_overD, r, v, n; };   // r = sqrt(area / π)
    // Log-space form: log m = la + b log D and log A (analytic except in the partially rimed regime, where the area is a sum)
    // give r = exp(log A / 2 - log π / 2) and the aspect-ratio factor ϕ^(1/3) = exp((log(3 sqrt π / 4 ρ) + log m - 3/2 log A) / 3),
    // which is folded into the exponents of the two velocity terms: 4 exp_ + 1 log per node instead of 5 exp_ + log + sqrt + rcp +
    // cbrt (the self-collection double integral evaluates 2048 nodes per point).  Same quantities, rounding-level differences.
    template <bool NEED_V, bool NEED_MELT = false, bool NEED_MASS = false>
    CM_DEV Node node(double D, const P3K& k) const {
        Node o;
        const double lD = logp_(D);
        const int r = regime(D, k);
        const double la = (r == 0) ? la_small : ((r == 3) ? la_grp : ((r == 4) ? la_part : la_unr));
        const double b = (r == 0 || r == 3) ? 3.0 : k.beta_va;
        const double lm = fma_(b, lD, la);
        o.mass = NEED_MASS ? exp_(lm) : 0.0;
        // ∂m/∂D / D = a b D^(b-2)
        o.dmass_dD_overD = NEED_MELT ? b * exp_(fma_(b - 2.0, lD, la)) : 0.0;
        const bool sphere = (r == 0 || r == 3);
        double larea = sphere ? fma_(2.0, lD, k.log_pi4) : fma_(k.sigma_a, lD, k.log_gamma_a);
        if (r == 4) {   // F_rim sphere + (1 - F_rim) non-spherical                    P3_particle_properties.jl:407-420
            const double sph = D * D * (num<double>::pi() / 4.0);
            const double non = exp_(larea);
            larea = logp_(F_rim * sph + (1.0 - F_rim) * non);
        }
        o.r = exp_(fma_(0.5, larea, -k.half_log_pi));
        o.n = exp_(fmax_(logN0 + mu * lD - lam * D, -700.0));
        if (NEED_V) {
            const bool small_ = D <= k.cutoff;
            const double A0 = small_ ? sa0 : ga0, B0 = small_ ? sb : gb0, A1 = small_ ? sa1 : ga1, B1 = small_ ? sb : gb1,
                         C1 = small_ ? sc1 : gc1;
            // ϕ = 3 sqrt(π) m / (4 ρ A^(3/2)), v *= ϕ^(1/3)                         P3_terminal_velocity.jl:32-60
            const double lp3 = k.aspect_oblate ? (((r == 3) ? lphc_g : k.lphc_i) + lm - 1.5 * larea) * (1.0 / 3.0) : 0.0;
            o.v = A0 * exp_(fma_(B0, lD, lp3)) + A1 * exp_(fmax_(fma_(B1, lD, -C1 * D), -700.0) + lp3);
        } else {
            o.v = 0.0;
        }
        return o;
    }
    CM_DEV double v_liq(double D, double lD) const {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) v += ra[j] * exp_(fmax_(fma_(rb[j], lD, -rc[j] * D), -700.0));
        return v;
    }
    CM_DEV double v_liq(double D) const { return v_liq(D, logp_(D)); }
    CM_DEV double v_liq_nl(double D, double lD) const {   // the same through the out-of-line exp (set-up code)
        double v = 0.0;
#pragma unroll 1
        for (int j = 0; j < 3; ++j) v += ra[j] * exp_nl_(fmax_(fma_(rb[j], lD, -rc[j] * D), -700.0));
        return v;
    }
};

// get_μ from λ = exp(logλ) (P3_size_distribution.jl:171-199): ONE definition, used by p3_point_init and by the quantile
// solves ahead of the point's pass from logλ — the same expressions, so a quantile solved ahead of the point's pass has the same bits
CM_DEV void p3_mu_lam(const P3K& k, double logl, double& mu, double& lam) {
    lam = exp_nl_(logl);
    mu = k.slope_power_law ? clamp_(k.slope_a * pow_pos_(lam, k.slope_b) - k.slope_c, 0.0, k.mu_max) : k.mu_const;
}
// P3.state_from_prognostic + P3State + PSD parameters + velocity coefficients
CM_DEV void p3_point_init(P3Point& s, const cumicro_params_p3_f64& p, const P3K& k, double rho, double T, double L_ice, double N_ice,
                          double L_rim, double B_rim, double logl) {
    s.rho = rho; s.T = T; s.L_ice = L_ice; s.N_ice = N_ice;
    s.F_rim = fmin_(regularised_ratio_(fmin_(L_rim, L_ice), L_ice, k.eps), 1.0 - k.eps);
    s.rho_rim = fmin_(regularised_ratio_(L_rim, B_rim, k.eps), k.rho_l08);
    // get_ρ_d (exprel form)                                            P3_particle_properties.jl:191-199
    const double pp = k.thr_p;
    const double logFu = log1p_nl_(-s.F_rim);
    auto exprel1 = [](double x) { return expm1_nl_(x) / x; };
    auto exprel2 = [](double x) {
        if (fabs(x) < 0.2) {
            double r = 1.0 / 362880.0;
            const double c[7] = {1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 1.0 / 2.0};
#pragma unroll
            for (int i = 0; i < 7; ++i) r = r * x + c[i];
            return r;
        }
        return (expm1_nl_(x) - x) / (x * x);
    };
    const double phi1 = exprel1(logFu);
    const double phi1mp = exprel1((1.0 - pp) * logFu);
    const double H = -pp * exprel2(-pp * logFu) - (1.0 - pp) * exprel2((1.0 - pp) * logFu);
    const double G = H - phi1mp * phi1;
    const double rho_d = -(s.rho_rim * phi1 * phi1mp) / G;
    s.rho_g = s.F_rim * s.rho_rim + (1.0 - s.F_rim) * rho_d;
    const bool unrimed = (s.F_rim == 0.0);
    const double pi = num<double>::pi();
    s.D_gr = unrimed ? num<double>::inf() : pow_pos_(k.thr_coef / (pi * s.rho_g), pp);
    s.D_cr = unrimed ? num<double>::inf() : pow_pos_(k.thr_coef / (pi * (s.rho_g * (1.0 - s.F_rim))), pp);
    // ice_mass_coeffs                                                 P3_particle_properties.jl:346-359
    const double Fu = fmax_(1.0 - s.F_rim, k.eps);
    s.la_small = k.log_rho_i_pi6;
    s.la_unr = k.log_alpha_va;
    s.la_grp = unrimed ? 0.0 : logp_nl_(s.rho_g * pi / 6.0);
    s.lphc_g = unrimed ? 0.0 : logp_nl_(k.phi_coef / (4.0 * s.rho_g));
    s.la_part = logp_nl_(k.alpha_va / Fu);
    // get_μ, get_logN₀                                               P3_size_distribution.jl:171-237
    p3_mu_lam(k, logl, s.mu, s.lam);
    {
        const double z = 0.0 + s.mu + 1.0;
        s.logN0 = logp_nl_(N_ice) - (-z * logl + lgamma_pos_(z) + 0.0);
    }
    // Chen 2022 coefficients at ρₐ                                     CO:290-349
    const double ra_ = fmax_(rho, 0.0);
    const double log1000 = 6.907755278982137;
    {
        const double pa = pow_pos_(ra_, k.As);
        const double b = k.Bs + ra_ * k.Cs;
        const double u = exp_nl_(b * log1000);
        s.sa0 = (k.Es * pa) * u; s.sa1 = (k.Fs * pa) * u; s.sb = b; s.sc1 = k.Gs1000;
        const double pl = pow_pos_(ra_, k.Al);
        s.ga0 = (k.Bl * pl) * k.pow1000_Cl;
        s.ga1 = (k.El * pl * exp_nl_(k.Hl * ra_)) * k.pow1000_Fl;
        s.gb0 = k.Cl; s.gb1 = k.Fl; s.gc1 = k.Gl1000;
    }
    {   // CO.Chen2022_vel_coeffs(::Chen2022VelTypeRain, ρₐ)                          CO:290-300
        const auto& v = p.vel_rain;
        const double q = exp_nl_(v.rho0 * ra_);
        const double ai[3] = {v.a[0] * q, v.a[1] * q, v.a[2] * q * pow_pos_(ra_, v.a3_pow)};
#pragma unroll 1
        for (int i = 0; i < 3; ++i) {
            s.rb[i] = v.b[i] - v.b_rho * ra_;
            s.ra[i] = ai[i] * exp_nl_(s.rb[i] * log1000);
            s.rc[i] = v.c[i] * 1000.0;
        }
    }
}

// P3.integral_bounds -> segment_boundaries                       P3_integral_properties.jl:34-45
CM_DEV void p3_segments(const P3Point& s, const P3K& k, double D_min, double D_max, double (&b)[5]) {
    b[0] = D_min;
    b[1] = clamp_(k.D_th, D_min, D_max);
    b[2] = clamp_(s.D_gr, D_min, D_max);
    b[3] = clamp_(s.D_cr, D_min, D_max);
    b[4] = D_max;
}

// Enumerates the quadrature nodes of the non-empty segments of b[0..4] across lanes:
// node index t in [0, count) -> (x, weight * scale).  Quadrature.integrate   src/Quadrature.jl:62-125
struct SegNodes {
    double lo[4], hw[4];   // shift, scale of the active segments
    int nact, n;
    CM_DEV void init(const double (&b)[5], int n_) {
        n = n_;
        nact = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (b[i] < b[i + 1]) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j == nact) { lo[j] = (b[i] + b[i + 1]) / 2.0; hw[j] = (b[i + 1] - b[i]) / 2.0; }
                ++nact;
            }
    }
    CM_DEV int count() const { return nact * n; }
    CM_DEV void get(int t, const double* qx, const double* qw, double& x, double& w) const {
        const int sgm = t / n, i = t - sgm * n;
        double shift = lo[0], scale = hw[0];
#pragma unroll
        for (int j = 1; j < 4; ++j)
            if (sgm == j) { shift = lo[j]; scale = hw[j]; }
        x = scale * qx[i] + shift;
        w = qw[i] * scale;
    }
};

// LocalRimeDensity callable                                        CMP/MicrophysicsP3.jl:222-239
CM_DEV double local_rime_density(const P3K& k, double Ri) {
    Ri = clamp_(Ri, 1.0, 12.0);
    if (Ri <= 8.0) return k.rim_a + k.rim_b * Ri + k.rim_c * (Ri * Ri);
    const double f = (Ri - 8.0) * 0.25;   // (Rᵢ - 8) / (12 - 8), exact
    return (1.0 - f) * k.rim8 + f * k.rim_rho_ice;
}

// RootSolvers.BrentsMethod with an always-false tolerance: Brent (1973), fixed iterations.
// Same statement order as the oracle's restatement (oracle/oracle_p3.hpp brent_fixed).
template <class F> CM_DEV double brent_fixed(F f, double a, double b, double fa, double fb, int iters) {
    if (fabs(fa) < fabs(fb)) { double t = a; a = b; b = t; t = fa; fa = fb; fb = t; }
    double c = a, fc = fa, d = c;
    bool mflag = true;
    for (int it = 0; it < iters; ++it) {
        if (fb == 0.0) break;
        double s;
        if (fa != fc && fb != fc)
            s = a * fb * fc / ((fa - fb) * (fa - fc)) + b * fa * fc / ((fb - fa) * (fb - fc)) + c * fa * fb / ((fc - fa) * (fc - fb));
        else
            s = b - fb * (b - a) / (fb - fa);
        double lo = (3.0 * a + b) / 4.0, hi = b;
        if (lo > hi) { const double t = lo; lo = hi; hi = t; }
        const bool bis = !(s > lo && s < hi) || (mflag && fabs(s - b) >= fabs(b - c) / 2.0) || (!mflag && fabs(s - b) >= fabs(c - d) / 2.0);
        if (bis) { s = (a + b) / 2.0; mflag = true; } else mflag = false;
        const double fs = f(s);
        d = c; c = b; fc = fb;
        if (fa * fs < 0.0) { b = s; fb = fs; } else { a = s; fa = fs; }
        if (fabs(fa) < fabs(fb)) { double t = a; a = b; b = t; t = fa; fa = fb; fb = t; }
    }
    return b;
}

struct P3Rates {
    double v_n, v_m;                 // ice_terminal_velocity_{number,mass}_weighted
    double melt_dN, melt_dL;         // ice_melt
    double agg_dN;                   // ice_self_collection
    double src[7];                   // bulk_liquid_ice_collision_sources: dq_c, dq_r, dN_c, dN_r, dL_rim, dL_ice, dB_rim
};

enum { P3_WANT_VEL = 1, P3_WANT_MELT = 2, P3_WANT_AGG = 4, P3_WANT_COLL = 8 };

// Phase barriers (kernels_p3.cu, CUMICRO_P3_SYNC >= 2; measured slower than the per-point barrier alone, kept for the record:
// 2^19 points, 1024x1: no barrier 64.5 ms, per-point barrier 50.7 ms, per-point + phase barriers 56.5 ms): every warp of the block executes exactly p3_phase_barriers(n) block
// barriers per point, at the same phase boundaries, so that the warps of an SM walk the same loops at the same time and
// share their instruction-cache lines.  All barrier sites are warp-uniform (`want`, `rain_on`, ... are broadcast values);
// a warp without a point (or without a phase) executes the matching barriers of the branch it skips.
#ifndef CUMICRO_P3_SYNC
#define CUMICRO_P3_SYNC 1
#endif
#if CUMICRO_P3_SYNC >= 2 && defined(__CUDA_ARCH__)
#define P3_BAR() do { __syncwarp(); asm volatile("bar.sync 0;" ::: "memory"); } while (0)
#else
#define P3_BAR() do { } while (0)
#endif
CM_HD int p3_coll_passes(int n) { return (4 * n + 31) / 32; }            // outer collision nodes: <= 4 segments x n, 32 per pass
CM_HD int p3_phase_barriers(int n) { return 4 + p3_coll_passes(n); }

// The quantile pairs the requested integrals need (one Halley solve per lane).
// Quantile `which` (0,1: p = 1e-6 and 1 - p: velocities, melt   2,3: p = 1e-5: collisions   4,5: p = eps: self-collection) of the
// ice PSD with slope exp(logl), as a diameter; 0 when the point's `want` mask does not need it.     P3_size_distribution.jl:171-237
CM_DEV double p3_quantile(const P3K& k, int want, int which, double mu, double lam) {
    const int g = which >> 1;
    const bool need = (g == 0) ? (want & (P3_WANT_VEL | P3_WANT_MELT)) : ((g == 1) ? (want & P3_WANT_COLL) : (want & P3_WANT_AGG));
    if (!need) return 0.0;
    const double pq = (g == 0) ? 1e-6 : ((g == 1) ? 0.00001 : k.eps);
    const double Y = (which & 1) ? (1.0 - pq) : pq;
    return gamma_inc_inv_(mu + 1.0, Y, 1.0 - Y, k.gamma_iters, k.eps) / lam;
}

// `have_pre`: lane l < 6 holds quantile l in `pre`, solved ahead (one per thread of the block, kernels_p3.cu); else one Halley
// solve per lane here.
CM_DEV void p3_bounds(const P3Point& s, const P3K& k, int want, int lane, double (&bv)[5], double (&bc)[5], double (&ba)[5],
                      bool have_pre = false, double pre = 0.0) {
    double x = 0.0;
    if (lane < 6) x = have_pre ? pre : p3_quantile(k, want, lane, s.mu, s.lam);
    p3_segments(s, k, bcast(x, 0), bcast(x, 1), bv);
    p3_segments(s, k, bcast(x, 2), bcast(x, 3), bc);
    p3_segments(s, k, bcast(x, 4), bcast(x, 5), ba);
}

// All requested P3 process rates of one point.  `qx`, `qw`: quadrature nodes / weights in shared
// memory; `sc`: this warp's scratch.  L_c, N_c, L_r, N_r: volumetric liquid contents.
CM_DEV void p3_point_rates(const P3Point& s, const cumicro_params_p3_f64& p, const P3K& k, const ThermoK<double>& tk,
                           const SB2006K<double>& sk, const double* qx, const double* qw, P3Scratch& sc, int want, double L_c,
                           double N_c, double L_r, double N_r, P3Rates& out, bool have_pre = false, double pre = 0.0) {
    const int lane = threadIdx.x & 31;
    const int n = k.n;
    const double pi = num<double>::pi();
    double bv[5], bc[5], ba[5];
    p3_bounds(s, k, want, lane, bv, bc, ba, have_pre, pre);
    SegNodes sn;
    P3_BAR();   // 1

    // ---- bulk terminal velocities + melt: single integrals over the p = 1e-6 bounds
    //      P3_terminal_velocity.jl:73-133, P3_processes.jl:64-94
    out.v_n = out.v_m = out.melt_dN = out.melt_dL = 0.0;
    if (want & (P3_WANT_VEL | P3_WANT_MELT)) {
        sn.init(bv, n);
        double a_n = 0.0, a_m = 0.0, a_melt = 0.0;
        for (int t = lane; t < sn.count(); t += 32) {
            double D, w;
            sn.get(t, qx, qw, D, w);
            const P3Point::Node nd = s.node<true, true, true>(D, k);
            const double nv = nd.n * nd.v;
            a_n += nv * w;
            a_m += nv * nd.mass * w;
            const double Fv = k.vent_a + k.vent_b * k.cbrt_Nsc * sqrt_(D * nd.v * k.inv_nu_air);
            a_melt += nd.dmass_dD_overD * Fv * nd.n * w;
        }
        const bool empty = (s.N_ice < k.eps) || (s.L_ice < k.eps);
        if (want & P3_WANT_VEL) {
            out.v_n = empty ? 0.0 : warp_sum(a_n) / s.N_ice;
            out.v_m = empty ? 0.0 : warp_sum(a_m) / s.L_ice;
        }
        if ((want & P3_WANT_MELT) && s.T > tk.T_freeze) {
            const double L_f = latent_heat_fusion(tk, s.T);
            const double fac = 4.0 * k.K_therm / L_f * (s.T - k.T_freeze);
            out.melt_dL = clamp0_(fac * warp_sum(a_melt));
            out.melt_dN = s.N_ice / s.L_ice * out.melt_dL;
        }
    }

    P3_BAR();   // 2
    // ---- ice self-collection: outer nodes by chunks of 32 (one per lane), inner 2n nodes across lanes
    //      P3_processes.jl:676-712
    out.agg_dN = 0.0;
    if (want & P3_WANT_AGG) {
        sn.init(ba, n);
        const double D_lo = ba[0], D_hi = ba[4];
        double acc = 0.0;
        const int tot = sn.count();
        for (int base = 0; base < tot; base += 32) {
            double D1 = 0.0, v1 = 0.0, r1 = 0.0, W1 = 0.0;
            if (base + lane < tot) {
                double w;
                sn.get(base + lane, qx, qw, D1, w);
                const P3Point::Node nd = s.node<true>(D1, k);
                v1 = nd.v;
                r1 = nd.r;
                W1 = nd.n * w;
            }
            const int cnt = min(32, tot - base);
            for (int t = 0; t < cnt; ++t) {
                const double d1 = bcast(D1, t), vv1 = bcast(v1, t), rr1 = bcast(r1, t), ww1 = bcast(W1, t);
                double part = 0.0;
                for (int u = lane; u < 2 * n; u += 32) {
                    const int half = (u >= n) ? 1 : 0;
                    const int i = u - half * n;
                    const double a = half ? d1 : D_lo, b = half ? D_hi : d1;
                    if (a < b) {
                        const double scale = (b - a) / 2.0, shift = (a + b) / 2.0;
                        const double D2 = scale * qx[i] + shift;
                        const P3Point::Node nd = s.node<true>(D2, k);
                        const double r2 = nd.r;
                        const double K = pi * ((rr1 + r2) * (rr1 + r2));
                        part += K * fabs(vv1 - nd.v) * nd.n * (qw[i] * scale);
                    }
                }
                acc += part * ww1;
            }
        }
        out.agg_dN = 0.5 * warp_sum(acc);
    }

    P3_BAR();   // 3
    // ---- liquid-ice collisions                                          P3_processes.jl:112-655
#pragma unroll
    for (int i = 0; i < 7; ++i) out.src[i] = 0.0;
    if (!(want & P3_WANT_COLL)) {
        for (int b = 0; b < 1 + p3_coll_passes(n); ++b) P3_BAR();
    } else {
        const double e = k.eps;
        const double rho = s.rho;
        // cloud PSD n_c(D) = exp(logN0c + νcD log D - λc D^μcD) and its p-quantile bounds   CM2:203-236, 346-355
        const double q_c = L_c / rho, q_r = L_r / rho;
        const bool cloud_off = (N_c < e) || (q_c < e);
        double logN0c, loglam_c, cb0, cb1;
        {
            const auto& pc = p.warm.sb.pdf_c;
            const double safe_q = fmax_(q_c, e), safe_N = fmax_(N_c, e);
            const double logx = logp_nl_(rho * safe_q / safe_N);
            const double z1 = (pc.nu_c + 1.0) / pc.mu_c;
            const double lB = -pc.mu_c * (logx + pc.loggamma_z1 - pc.loggamma_z2);
            const double lA = k.log_mu_c + logp_nl_(safe_N) + z1 * lB - pc.loggamma_z1;
            logN0c = lA + k.log3 + (pc.nu_c + 1.0) * k.log_km;
            loglam_c = lB + pc.mu_c * k.log_km;
            cb0 = exp_nl_((k.cloud_log_z_lo - loglam_c) / k.mu_cD);
            cb1 = exp_nl_((k.cloud_log_z_hi - loglam_c) / k.mu_cD);
        }
        // rain PSD, bounds                                                             CM2:270-276, 337-345
        const RainPDF<double> rp = pdf_rain_parameters<double>(p.warm.sb.pdf_r, sk.pi_rho_w, e, q_r, rho, N_r);
        double rb0 = 0.0, rb1 = 0.0;
        if (!(rp.Dr_mean == 0.0)) {
            const double lDr = logp_nl_(rp.Dr_mean);
            rb0 = exp_nl_(lDr + k.rain_cll_lo);
            rb1 = exp_nl_(lDr + k.rain_cll_hi);
        }
        const bool rain_on = !(rp.N0r == 0.0 || !(rb1 > rb0));
        const bool cloud_on = !cloud_off && (cb0 < cb1);
        const double TC = s.T - k.T_freeze;
        const double inv_two_TC = 1.0 / (2.0 * TC);   // Inf at T = T_freeze, like the reference's division by zero

        // inner nodes -> shared
        __syncwarp();
        for (int j = lane; j < n; j += 32) {
            {
                const double scale = (cb1 - cb0) / 2.0, shift = (cb0 + cb1) / 2.0;
                const double D = cloud_on ? scale * qx[j] + shift : 1e-6;
                const double lD = logp_nl_(D);
                const double nc = exp_nl_(logN0c + k.nu_cD * lD - exp_nl_(fma_(k.mu_cD, lD, loglam_c)));
                sc.cD[j] = D;
                sc.cWN[j] = cloud_on ? nc * (qw[j] * scale) : 0.0;
                sc.cM[j] = k.mliq_coef * (D * D * D * pi / 6.0);
                sc.cV[j] = s.v_liq_nl(D, lD);
            }
            {
                const double scale = (rb1 - rb0) / 2.0, shift = (rb0 + rb1) / 2.0;
                const double D = rain_on ? scale * qx[j] + shift : 1e-6;
                const double lD = logp_nl_(D);
                const double nr = rp.N0r * exp_nl_(-D / (rain_on ? rp.Dr_mean : 1.0));
                sc.rD[j] = D;
                sc.rWNM[j] = rain_on ? nr * (k.mliq_coef * (D * D * D * pi / 6.0)) * (qw[j] * scale) : 0.0;
                sc.rV[j] = s.v_liq_nl(D, lD);
            }
        }
        // closed-form rain inner integral: the (z, α) table at the fixed ends       P3_processes.jl:344-369
        const double lam_r = rain_on ? 1.0 / rp.Dr_mean : 1.0;
        double vl_min = 0.0, vl_max = 0.0;
        if (rain_on) {
            vl_min = s.v_liq_nl(rb0, logp_nl_(rb0));
            vl_max = s.v_liq_nl(rb1, logp_nl_(rb1));
            if (lane < kGam) {
                const int j = lane / 6, pi_ = lane - j * 6;     // velocity term (0: the v_i term), p + i
                const double cj = (j == 0) ? 0.0 : s.rc[j - 1];
                const double alpha = lam_r + cj;
                const double p0v = (pi_ >= 3) ? 3.0 : 0.0;
                const double pj = (j == 0) ? p0v : p0v + s.rb[j - 1];     // flux: p + bi[j]
                const double z = (pj + (double)(pi_ - (int)p0v)) + 1.0;   // Iᵖ: p + (i - 1); gamma_inc_moment: z = p + 1
                const double tg = tgamma_pos_(z);
                const double lg = logp_nl_(tg);
                sc.gz[lane] = z;
                sc.glg[lane] = lg;
                sc.gG[lane] = tg / pow_pos_(alpha, z);
                PQ g = gamma_inc_(z, alpha * rb0, lg, k.gamma_iters);
                sc.gP0[lane] = g.P; sc.gQ0[lane] = g.Q;
                g = gamma_inc_(z, alpha * rb1, lg, k.gamma_iters);
                sc.gP1[lane] = g.P; sc.gQ1[lane] = g.Q;
                if (pi_ == 0) { sc.tA[j] = alpha; sc.tC[j] = (j == 0) ? 0.0 : s.ra[j - 1]; }
            }
        }
        __syncwarp();

        // compute_max_freeze_rate: the D-independent part                          P3_processes.jl:184-219
        const double T_frz = tk.T_freeze;
        const double dT = T_frz - s.T;
        const TempState<double> ts_f = temp_state(tk, T_frz), ts_a = temp_state(tk, s.T);
        const double drho_sat = rho * (p_sat_ice(tk, ts_f) / (tk.R_v * rho * T_frz) - p_sat_ice(tk, ts_a) / (tk.R_v * rho * s.T));
        const double Lv = latent_heat_vapor(tk, s.T), L_f = latent_heat_fusion(tk, s.T);
        const double denom = L_f - tk.cv_l * dT;
        const double frz_num = k.K_therm * dT + Lv * k.D_vapor * drho_sat;
        const double mfac = k.rho_w * (1.0 * 1.0 * 1.0 * pi / 6.0);

        sn.init(bc, n);
        double acc[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) acc[i] = 0.0;
        P3_BAR();   // 4
#if CUMICRO_P3_SYNC >= 2
        for (int t = lane, pass = 0; pass < p3_coll_passes(n); t += 32, ++pass) {
            P3_BAR();   // 5 ...
            if (t >= sn.count()) continue;
#else
        for (int t = lane; t < sn.count(); t += 32) {
#endif
            double Di, w;
            sn.get(t, qx, qw, Di, w);
            const P3Point::Node nd = s.node<true>(Di, k);
            const double v_i = nd.v;
            const double r_i = nd.r;
            const double k0 = pi * (r_i * r_i), k1 = pi * r_i, k2 = pi / 4.0;
            // cloud inner integrals (N, M, B)                                            :304-319
            double cN = 0.0, cMm = 0.0, cB = 0.0;
#pragma unroll 1
            for (int j = 0; j < n; ++j) {
                const double Dl = sc.cD[j];
                const double dv = fabs(v_i - sc.cV[j]);
                const double K = k0 + Dl * (k1 + Dl * k2);
                const double t1 = (K * dv) * sc.cWN[j];
                const double t2 = t1 * sc.cM[j];
                const double Ri = (Dl * 1000000.0 * dv) * inv_two_TC;
                cN += t1;
                cMm += t2;
                cB += t2 * rcp_(local_rime_density(k, Ri));
            }
            // rain inner integrals: closed form for N, M; quadrature for B             :381-415
            double rN = 0.0, rM = 0.0, rB = 0.0;
            if (rain_on) {
                const double Dstar = brent_fixed([&](double D) { return s.v_liq(D) - v_i; }, rb0, rb1, vl_min - v_i, vl_max - v_i, k.brent_iters);
                // flux(a, b, p) = v_i Iᵖ(a, b, p, λ) - Σ_j a_j Iᵖ(a, b, p + b_j, λ + c_j) over the two pieces, p = 0 | 3
                double f_lo0 = 0.0, f_hi0 = 0.0, f_lo3 = 0.0, f_hi3 = 0.0;
#pragma unroll 1
                for (int j = 0; j < 4; ++j) {
                    const double alpha = sc.tA[j];
                    const double xs = alpha * Dstar, x1 = alpha * rb1;
                    const double lxs = (xs > 0.0 && xs < num<double>::inf()) ? logp_nl_(xs) : 0.0;   // one logarithm for the six orders z
                    double a_lo0 = 0.0, a_hi0 = 0.0, a_lo3 = 0.0, a_hi3 = 0.0;
                    double fac = 0.0, z_prev = 1.0;
#pragma unroll 1
                    for (int pi_ = 0; pi_ < 6; ++pi_) {
                        const int g = j * 6 + pi_;
                        const double z = sc.gz[g];                    // six consecutive orders: z_0, z_0 + 1, ..., z_0 + 5
                        fac = (pi_ == 0) ? gamma_inc_factor_(z, xs, lxs, sc.glg[g]) : fac * (xs * rcp_(z_prev));
                        z_prev = z;
                        const PQ q = gamma_inc_core_(z, xs, fac, k.gamma_iters);
                        // gamma_inc_moment(D_min, Dstar) and (Dstar, D_max)                P3_size_distribution.jl:121-133
                        double m_lo = 0.0, m_hi = 0.0;
                        if (Dstar > rb0) m_lo = sc.gG[g] * fmax_((xs < z + 1.0) ? q.P - sc.gP0[g] : sc.gQ0[g] - q.Q, 0.0);
                        if (rb1 > Dstar) m_hi = sc.gG[g] * fmax_((x1 < z + 1.0) ? sc.gP1[g] - q.P : q.Q - sc.gQ1[g], 0.0);
                        const int i = (pi_ >= 3) ? pi_ - 3 : pi_;
                        const double coef = (i == 0) ? k0 : ((i == 1) ? k1 : k2);
                        const double t_lo = coef * m_lo, t_hi = coef * m_hi;
                        if (pi_ < 3) { a_lo0 = (i == 0) ? t_lo : a_lo0 + t_lo; a_hi0 = (i == 0) ? t_hi : a_hi0 + t_hi; }
                        else         { a_lo3 = (i == 0) ? t_lo : a_lo3 + t_lo; a_hi3 = (i == 0) ? t_hi : a_hi3 + t_hi; }
                    }
                    if (j == 0) { f_lo0 = v_i * a_lo0; f_hi0 = v_i * a_hi0; f_lo3 = v_i * a_lo3; f_hi3 = v_i * a_hi3; }
                    else {
                        const double aj = sc.tC[j];
                        f_lo0 -= aj * a_lo0; f_hi0 -= aj * a_hi0; f_lo3 -= aj * a_lo3; f_hi3 -= aj * a_hi3;
                    }
                }
                const double dN = rp.N0r * (f_lo0 - f_hi0);
                const double dM = rp.N0r * mfac * (f_lo3 - f_hi3);
                if (isfinite(dN) && isfinite(dM)) {
                    rN = dN;
                    rM = dM;
#pragma unroll 1
                    for (int j = 0; j < n; ++j) {
                        const double Dl = sc.rD[j];
                        const double dv = fabs(v_i - sc.rV[j]);
                        const double K = k0 + Dl * (k1 + Dl * k2);
                        const double Ri = (Dl * 1000000.0 * dv) * inv_two_TC;
                        rB += (K * dv) * sc.rWNM[j] * rcp_(local_rime_density(k, Ri));
                    }
                }
            }
            // partition between freezing and shedding                                   :462-489
            const double M_col = cMm + rM;
            double M_max;
            if (s.T >= T_frz) M_max = 0.0;
            else if (!(denom > 0.0)) M_max = 1.7976931348623157e308;
            else {
                const double Fv = k.vent_a + k.vent_b * k.cbrt_Nsc * sqrt_(Di * v_i * k.inv_nu_air);
                M_max = 2.0 * (pi * Di) * Fv * frz_num / denom;
            }
            const double M_frz = fmin_(M_col, M_max);
            const double f_frz = (M_col == 0.0) ? 0.0 : M_frz / M_col;
            const double wet = (M_col > M_frz) ? 1.0 : 0.0;
            const double nw = nd.n * w;
            acc[0] += nw * cMm * f_frz;
            acc[1] += nw * cMm * (1.0 - f_frz);
            acc[2] += nw * cN;
            acc[3] += nw * rM * f_frz;
            acc[4] += nw * rM * (1.0 - f_frz);
            acc[5] += nw * rN;
            acc[6] += nw * M_col;
            acc[7] += nw * cB * f_frz;
            acc[8] += nw * rB * f_frz;
            acc[9] += nw * wet * M_col;
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) acc[i] = warp_sum(acc[i]);
        // bulk_liquid_ice_collision_sources                                              :606-655
        const double QCFRZ = acc[0], QCSHD = acc[1], NCCOL = acc[2], QRFRZ = acc[3], QRSHD = acc[4], NRCOL = acc[5], M_col = acc[6],
                     BCCOL = acc[7], BRCOL = acc[8], wetM = acc[9];
        const double f_wet = (M_col == 0.0) ? 0.0 : wetM / M_col;
        const double NRSHD = QRSHD / k.m_shd;
        const double B_rim = (s.rho_rim == 0.0) ? 0.0 : (s.L_ice * s.F_rim) / s.rho_rim;
        const double QIWET = f_wet * s.L_ice * (1.0 - s.F_rim) / k.tau_wet;
        const double BIWET = f_wet * (s.L_ice / k.rho_i - B_rim) / k.tau_wet;
        out.src[0] = (-QCFRZ - QCSHD) / rho;
        out.src[1] = (-QRFRZ + QCSHD) / rho;
        out.src[2] = -NCCOL;
        out.src[3] = -NRCOL + NRSHD;
        out.src[4] = QCFRZ + QRFRZ + QIWET;
        out.src[5] = QCFRZ + QRFRZ;
        out.src[6] = BCCOL + BRCOL + BIWET;
    }
}
#endif  // __CUDACC__

}  // namespace cm
