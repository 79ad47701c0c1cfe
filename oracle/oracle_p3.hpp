// oracle_p3.hpp — CPU restatement of the P3 ice scheme: src/P3_particle_properties.jl,
// P3_size_distribution.jl, P3_integral_properties.jl, P3_terminal_velocity.jl, P3_processes.jl,
// src/Quadrature.jl, the incomplete-gamma utilities of src/Utilities.jl (UT:92-252, 399-509),
// src/DistributionTools.jl, the PSD closures of src/Microphysics2M.jl (CM2:203-355), the Bigg /
// Frostenberg-2023 rates of src/IceNucleation.jl (IN:274-526) and the 2-moment + P3 method of
// src/BulkMicrophysicsTendencies.jl (BMT:898-1083).
// TEST INFRASTRUCTURE ONLY (see oracle_base.hpp).  Operation order follows the Julia source.
//
// Third-party pieces restated from their published definitions (not under /root/reference):
//   RootSolvers.jl BrentsMethod with an always-false tolerance (fixed 10 / 8 iterations): Brent
//   (1973) bisection / secant / inverse-quadratic iteration.  The exact RootSolvers iterate after
//   a fixed number of steps is NOT pinned by any reference test; both uses on
//   this path are smooth monotone problems; the 10th iterate is converged to 1e-9 on > 90 % of the
//   reference's own sweep of states and within 0.05 in logλ on all (tests/test_oracle_p3.py).
//   PARITY UNPINNED for the Brent iterates: logλ is an input of every downstream P3 function, and
//   the crossover diameter only enters the rain inner integral at second order.
//   FastGaussQuadrature.gausslegendre(n): nodes / weights are computed host-side (numpy
//   leggauss, exact to 1 ulp) and arrive in the parameter block.
#pragma once
#include "oracle_icenuc.hpp"

namespace orc {

template <class FT> struct PT3 {
    using params = typename std::conditional<std::is_same<FT, float>::value, cumicro_params_p3_f32, cumicro_params_p3_f64>::type;
    using scheme = typename std::conditional<std::is_same<FT, float>::value, cumicro_p3_scheme_f32, cumicro_p3_scheme_f64>::type;
    using quad = typename std::conditional<std::is_same<FT, float>::value, cumicro_quadrature_f32, cumicro_quadrature_f64>::type;
};

// Iteration counts of the fixed-iteration algorithms follow the METHOD's float type: the Float32
// method (also when it is evaluated in Float64 arithmetic, f32_thresholds mode) uses 20 / 8.
template <class FT> inline bool is_float32() { return std::is_same<FT, float>::value || f32_thresholds(); }
template <class FT> inline FT eps_type() { return eps<FT>(); }
template <class FT> inline FT nan_() { return std::numeric_limits<FT>::quiet_NaN(); }
template <> inline Tr nan_<Tr>() { return Tr(std::numeric_limits<double>::quiet_NaN()); }
template <class FT> inline FT floatmax_() { return std::numeric_limits<FT>::max(); }
template <> inline Tr floatmax_<Tr>() { return Tr(std::numeric_limits<double>::max()); }
template <class FT> inline double val_(FT x) { return double(x); }
inline double val_(Tr x) { return x.v; }

// ---- UT.gamma_inc (fixed-iteration series / Lentz continued fraction)            UT:92-144
template <class FT> inline void gamma_inc(FT a, FT x, FT loggamma_a, FT& P, FT& Q) {
    if (x <= FT(0)) { P = FT(0); Q = FT(1); return; }
    if (isinf_(x)) { P = FT(1); Q = FT(0); return; }
    FT factor = exp_(a * log_(x) - x - loggamma_a);
    const int maxiters = is_float32<FT>() ? 20 : 30;
    if (x < a + 1) {
        FT term = FT(1) / a;
        FT sum = term;
        for (int k = 1; k <= maxiters; ++k) {
            term *= x / (a + k);
            sum += term;
        }
        P = jclamp(FT(factor * sum), FT(0), FT(1));
        Q = FT(1) - P;
    } else {
        const FT tiny = FT(1e-30);
        FT b_1 = x + 1 - a;
        FT c = b_1 + 1 / tiny;
        FT d = 1 / b_1;
        FT h = d;
        for (int k = 1; k <= maxiters; ++k) {
            FT a_k = -FT(k) * (FT(k) - a);
            FT b_k = x + 2 * k + 1 - a;
            FT d_tmp = b_k + a_k * d;
            d = (fabs_(d_tmp) < tiny) ? tiny : d_tmp;
            FT c_tmp = b_k + a_k / c;
            c = (fabs_(c_tmp) < tiny) ? tiny : c_tmp;
            d = 1 / d;
            FT delta = c * d;
            h *= delta;
        }
        Q = jclamp(FT(factor * h), FT(0), FT(1));
        P = FT(1) - Q;
    }
}
template <class FT> inline void gamma_inc(FT a, FT x, FT& P, FT& Q) { gamma_inc(a, x, FT(lgamma_(a)), P, Q); }

// ---- UT.gamma_inc_inv (Halley, <= 15 steps with the reference's early exits)     UT:205-252
template <class FT> inline FT gamma_inc_inv(FT a, FT p, FT q) {
    if (p <= FT(0)) return FT(0);
    if (q <= FT(0)) return inf<FT>();
    FT x = (p < FT(0.5)) ? FT(pow_(p * tgamma_(a + 1), 1 / a)) : FT(a - log_(q));
    const bool use_q = p > FT(0.5);
    FT lga = lgamma_(a);
    for (int i = 1; i <= 15; ++i) {
        FT P, Q;
        gamma_inc(a, x, lga, P, Q);
        FT f = use_q ? FT(Q - q) : FT(P - p);
        FT fprime = exp_((a - 1) * log_(x) - x - lga);
        fprime = use_q ? FT(-fprime) : fprime;
        if (fprime == FT(0)) break;
        FT f2 = (a - 1 - x) / x;
        FT step = f / (fprime * (FT(1) - FT(0.5) * f * f2 / fprime));
        if (x - step <= FT(0)) step = FT(0.5) * x;
        x = x - step;
        if (fabs_(step) < eps_type<FT>() * x) break;
    }
    return x;
}

// ---- DistributionTools.jl                                                      DT:44-191
template <class FT> inline FT generalized_gamma_quantile(FT nu, FT mu, FT B, FT Y) {
    FT z = gamma_inc_inv<FT>((nu + 1) / mu, Y, 1 - Y);
    return pow_(z / B, 1 / mu);
}
template <class FT> inline FT generalized_gamma_quantile_unit_mu(FT nu, FT B, FT Y) { return gamma_inc_inv<FT>(nu + 1, Y, 1 - Y) / B; }
template <class FT> inline FT exponential_quantile(FT D_mean, FT Y) { return exp_(log_(D_mean) + cloglog(Y)); }
template <class FT> inline FT exponential_Mn(FT D_mean, FT N, int n) { return N * FT(fac(n)) * pow_(D_mean, FT(n)); }

// ---- UT.sgs_weight_function / _regularised_ratio / rime_mass_fraction / rime_density   UT:445-509
template <class FT> inline FT sgs_weight_function(FT a, FT a_half) {
    if (a < FT(0)) return FT(0);
    if (a > jmin(FT(1), 42 * a_half)) return FT(1);
    if (4 * a < eps_type<FT>()) return FT(0);
    return (1 + tanh_(2 * atanh_(1 - 2 * pow_(1 - a, -1 / log2_(1 - a_half))))) / 2;
}
template <class FT> inline FT regularised_ratio(FT num, FT den) {
    const FT half = eps_type<FT>(), e2 = eps_type<FT>() * eps_type<FT>();
    FT weight = sgs_weight_function<FT>(den, half);
    return (den < e2) ? FT(0) : FT(weight * num / den);
}

// ---- P3State                                                  P3_particle_properties.jl:20-106
template <class FT> struct P3State {
    const typename PT3<double>::scheme* prm;   // parameters stay plain Float64 (or widened)
    FT L_ice, N_ice, F_rim, rho_rim, rho_g, D_th, D_gr, D_cr;
};
template <class FT> inline FT exprel1(FT x) { return expm1_(x) / x; }
template <class FT> inline FT exprel2(FT x) {
    if (fabs_(x) < FT(1.0 / 5)) {   // evalpoly(x, 1/(i+1)!, i = 1..8)
        FT r = FT(1.0 / fac(9));
        for (int i = 7; i >= 1; --i) r = r * x + FT(1.0 / fac(i + 1));
        return r;
    }
    return (expm1_(x) - x) / (x * x);
}
template <class FT, class S> inline FT get_rho_d(const S& prm, FT F_rim, FT rho_rim) {
    FT p = 1 / (3 - FT(prm.beta_va));
    FT logFu = log1p_(-F_rim);
    FT phi1 = exprel1<FT>(logFu);
    FT phi1mp = exprel1<FT>((1 - p) * logFu);
    FT H = -p * exprel2<FT>(-p * logFu) - (1 - p) * exprel2<FT>((1 - p) * logFu);
    FT G = H - phi1mp * phi1;
    return -(rho_rim * phi1 * phi1mp) / G;
}
template <class FT, class S> inline FT p3_threshold(const S& prm, FT rho) {
    return pow_(6 * FT(prm.alpha_va) / (pi<FT>() * rho), 1 / (3 - FT(prm.beta_va)));
}
template <class FT, class S> inline P3State<FT> make_p3_state(const S& prm, FT L_ice, FT N_ice, FT F_rim, FT rho_rim) {
    P3State<FT> s;
    s.prm = &prm;
    s.L_ice = L_ice; s.N_ice = N_ice; s.F_rim = F_rim; s.rho_rim = rho_rim;
    FT rho_d = get_rho_d<FT>(prm, F_rim, rho_rim);
    s.rho_g = F_rim * rho_rim + (1 - F_rim) * rho_d;                     // weighted_average
    s.D_th = p3_threshold<FT>(prm, FT(prm.rho_i));
    const bool unrimed = (F_rim == FT(0));
    s.D_gr = unrimed ? inf<FT>() : p3_threshold<FT>(prm, s.rho_g);
    s.D_cr = unrimed ? inf<FT>() : p3_threshold<FT>(prm, s.rho_g * (1 - F_rim));
    return s;
}
template <class FT, class S> inline P3State<FT> state_from_prognostic(const S& prm, FT L_ice, FT N_ice, FT L_rim, FT B_rim) {
    FT F_rim = jmin(regularised_ratio<FT>(jmin(L_rim, L_ice), L_ice), FT(1) - eps_type<FT>());
    FT rho_rim = jmin(regularised_ratio<FT>(L_rim, B_rim), FT(0.8) * FT(prm.rho_l));
    return make_p3_state<FT>(prm, L_ice, N_ice, F_rim, rho_rim);
}
template <class FT> inline void segment_boundaries(const P3State<FT>& s, FT D_min, FT D_max, FT b[5]) {
    b[0] = D_min; b[1] = jclamp(s.D_th, D_min, D_max); b[2] = jclamp(s.D_gr, D_min, D_max); b[3] = jclamp(s.D_cr, D_min, D_max); b[4] = D_max;
}
template <class FT> inline FT regime_value(const P3State<FT>& s, FT D, FT small, FT unrimed, FT dense, FT graupel, FT partial) {
    return (D < s.D_th) ? small : ((s.F_rim == FT(0)) ? unrimed : ((D < s.D_gr) ? dense : ((D < s.D_cr) ? graupel : partial)));
}
template <class FT> inline void ice_mass_coeffs(const P3State<FT>& s, FT D, FT& a, FT& b) {
    const auto& p = *s.prm;
    FT Fu = jmax(1 - s.F_rim, eps_type<FT>());
    FT av = FT(p.alpha_va), bv = FT(p.beta_va);
    a = regime_value<FT>(s, D, FT(p.rho_i) * pi<FT>() / 6, av, av, s.rho_g * pi<FT>() / 6, av / Fu);
    b = regime_value<FT>(s, D, FT(3), bv, bv, FT(3), bv);
}
template <class FT> inline FT ice_mass(const P3State<FT>& s, FT D) { FT a, b; ice_mass_coeffs(s, D, a, b); return a * pow_(D, b); }
template <class FT> inline FT dice_mass_dD(const P3State<FT>& s, FT D) { FT a, b; ice_mass_coeffs(s, D, a, b); return (a * b) * pow_(D, b - 1); }
template <class FT> inline FT ice_area(const P3State<FT>& s, FT D) {
    const auto& p = *s.prm;
    FT spherical = D * D * pi<FT>() / 4;
    FT nonspherical = FT(p.gamma) * pow_(D, FT(p.sigma));
    return regime_value<FT>(s, D, spherical, nonspherical, nonspherical, spherical, s.F_rim * spherical + (1 - s.F_rim) * nonspherical);
}
template <class FT> inline FT phi_i(const P3State<FT>& s, FT D) {
    const auto& p = *s.prm;
    FT m = ice_mass(s, D), a = ice_area(s, D);
    FT rho = regime_value<FT>(s, D, FT(p.rho_i), FT(p.rho_i), FT(p.rho_i), s.rho_g, FT(p.rho_i));
    FT ob = 3 * sqrt_(pi<FT>()) * m / (4 * rho * a * sqrt_(a));
    return (D == FT(0)) ? FT(0) : ob;
}

// ---- size distribution                                            P3_size_distribution.jl:8-237
template <class FT, class S> inline FT get_mu(const S& prm, FT logl) {
    if (!prm.slope_power_law) return FT(prm.slope_mu_const);
    return jclamp(FT(FT(prm.slope_a) * pow_(exp_(logl), FT(prm.slope_b)) - FT(prm.slope_c)), FT(0), FT(prm.slope_mu_max));
}
template <class FT> inline FT loggamma_moment(FT mu, FT logl, FT k, FT scale) {
    FT z = k + mu + 1;
    return -z * logl + lgamma_(z) + log_(scale);
}
template <class FT> inline FT loggamma_inc_moment(FT D1, FT D2, FT mu, FT logl, FT k, FT scale) {
    if (!(D1 < D2)) return log_(FT(0));
    FT z = k + mu + 1;
    FT x1 = D1 * exp_(logl), x2 = D2 * exp_(logl);    // LogExpFunctions.xexpy
    FT p1, q1, p2, q2;
    gamma_inc(z, x1, p1, q1);
    gamma_inc(z, x2, p2, q2);
    FT dq = (x2 < z + 1) ? FT(p2 - p1) : FT(q1 - q2);
    dq = jmax(dq, eps_type<FT>());
    return -z * logl + lgamma_(z) + log_(dq) + log_(scale);
}
template <class FT> inline FT gamma_inc_moment(FT D1, FT D2, FT p, FT alpha) {
    if (!(D2 > D1)) return FT(0);
    if (!(alpha > FT(0))) return nan_<FT>();
    FT z = p + 1;
    FT x1 = alpha * D1, x2 = alpha * D2;
    FT p1, q1, p2, q2;
    gamma_inc(z, x1, p1, q1);
    gamma_inc(z, x2, p2, q2);
    FT dq = (x2 < z + 1) ? FT(p2 - p1) : FT(q1 - q2);
    dq = jmax(dq, FT(0));
    return tgamma_(z) * dq / pow_(alpha, z);
}
template <class FT> inline FT logsumexp4(const FT x[4]) {            // UT.unrolled_logsumexp  UT:399-412
    FT m = x[0];
    for (int i = 1; i < 4; ++i) m = jmax(m, x[i]);
    if (!isfinite_(m)) return m;
    FT s = FT(0);
    for (int i = 0; i < 4; ++i) s += exp_(x[i] - m);
    return m + log_(s);
}
template <class FT> inline FT logmass_gamma_moment(const P3State<FT>& s, FT mu, FT logl, FT n) {
    FT b[5];
    segment_boundaries<FT>(s, FT(0), inf<FT>(), b);
    FT m[4];
    for (int i = 0; i < 4; ++i) {
        FT a_, b_;
        ice_mass_coeffs(s, FT((b[i] + b[i + 1]) / 2), a_, b_);
        m[i] = loggamma_inc_moment<FT>(b[i], b[i + 1], mu, logl, b_ + n, a_);
    }
    return logsumexp4(m);
}
template <class FT> inline FT logLdivN(const P3State<FT>& s, FT logl) {
    FT mu = get_mu<FT>(*s.prm, logl);
    return logmass_gamma_moment<FT>(s, mu, logl, FT(0)) - loggamma_moment<FT>(mu, logl, FT(0), FT(1));
}
template <class FT> inline FT get_logN0(FT N_ice, FT mu, FT logl) { return log_(N_ice) - loggamma_moment<FT>(mu, logl, FT(0), FT(1)); }
template <class FT> struct IcePSD {   // P3LogNumberFunctor / P3SizeDistributionFunctor
    FT logN0, mu, lam;
    FT operator()(FT D) const { return exp_(logN0 + mu * log_(D) - lam * D); }
};
template <class FT> inline IcePSD<FT> ice_psd(const P3State<FT>& s, FT logl) {
    IcePSD<FT> f;
    f.mu = get_mu<FT>(*s.prm, logl);
    f.logN0 = get_logN0<FT>(s.N_ice, f.mu, logl);
    f.lam = exp_(logl);
    return f;
}

// ---- Brent's method, fixed number of iterations (RootSolvers.BrentsMethod + FixedIterations)
template <class FT, class F> inline FT brent_fixed(F f, FT a, FT b, int maxiters) {
    FT fa = f(a), fb = f(b);
    if (fabs_(fa) < fabs_(fb)) { std::swap(a, b); std::swap(fa, fb); }
    FT c = a, fc = fa, d = c;
    bool mflag = true;
    for (int it = 0; it < maxiters; ++it) {
        if (fb == FT(0)) break;
        FT s;
        if (fa != fc && fb != fc)
            s = a * fb * fc / ((fa - fb) * (fa - fc)) + b * fa * fc / ((fb - fa) * (fb - fc)) + c * fa * fb / ((fc - fa) * (fc - fb));
        else
            s = b - fb * (b - a) / (fb - fa);
        FT lo = (3 * a + b) / 4, hi = b;
        if (lo > hi) std::swap(lo, hi);
        bool bis = !(s > lo && s < hi) || (mflag && fabs_(s - b) >= fabs_(b - c) / 2) || (!mflag && fabs_(s - b) >= fabs_(c - d) / 2);
        if (bis) { s = (a + b) / 2; mflag = true; } else mflag = false;
        FT fs = f(s);
        d = c; c = b; fc = fb;
        if (val_(fa) * val_(fs) < 0) { b = s; fb = fs; } else { a = s; fa = fs; }
        if (fabs_(fa) < fabs_(fb)) { std::swap(a, b); std::swap(fa, fb); }
    }
    return b;
}

// P3.get_distribution_logλ                                          P3_size_distribution.jl:284-326
template <class FT> inline FT get_distribution_loglambda(const P3State<FT>& s, int maxiters = -1) {
    if (s.N_ice < eps_type<FT>() || s.L_ice < eps_type<FT>()) return log_(FT(0));
    FT target = log_(s.L_ice) - log_(s.N_ice);
    auto sp = [&](FT l) { return FT(logLdivN<FT>(s, l) - target); };
    FT lo = FT(2), hi = FT(17);
    FT f_lo = sp(lo), f_hi = sp(hi);
    if (!isfinite_(f_lo) || !isfinite_(f_hi) || val_(f_lo) * val_(f_hi) > 0) return (fabs_(f_lo) <= fabs_(f_hi)) ? lo : hi;
    if (maxiters < 0) maxiters = is_float32<FT>() ? 8 : 10;
    return brent_fixed<FT>(sp, lo, hi, maxiters);
}

// P3.integral_bounds                                              P3_integral_properties.jl:34-45
template <class FT> inline void integral_bounds(const P3State<FT>& s, FT logl, FT p, FT b[5]) {
    FT k = get_mu<FT>(*s.prm, logl);
    FT lam = exp_(logl);
    FT D_min = generalized_gamma_quantile_unit_mu<FT>(k, lam, p);
    FT D_max = generalized_gamma_quantile_unit_mu<FT>(k, lam, FT(1 - p));
    segment_boundaries<FT>(s, D_min, D_max, b);
}
// P3.D_m
template <class FT> inline FT D_m(const P3State<FT>& s, FT logl) {
    FT mu = get_mu<FT>(*s.prm, logl);
    return exp_(get_logN0<FT>(s.N_ice, mu, logl) + logmass_gamma_moment<FT>(s, mu, logl, FT(1))) / s.L_ice;
}

// ---- Quadrature.integrate                                                src/Quadrature.jl:62-125
template <class FT, class Q, class F> inline auto integrate1(F f, FT a, FT b, const Q& quad) -> decltype(f(a)) {
    using R = decltype(f(a));
    FT scale = (b - a) / 2, shift = (a + b) / 2;
    R result = R{} * FT(0);
    if (!(a < b)) return result;
    const int n = quad.n;
    for (int i = 1; i <= n; ++i) {
        FT y, w;
        if (quad.gauss_legendre) {
            y = FT(quad.nodes[i - 1]);
            w = FT(quad.weights[i - 1]);
        } else {   // ChebyshevGauss: node cospi((2i-1)/(2n)), weight π/n, inverse weight function sqrt(1-y²)
            y = FT(std::cos(3.141592653589793238462643383279502884L * (2.0L * i - 1) / (2.0L * n)));
            w = sqrt_(1 - y * y) * (pi<FT>() / FT(n));
        }
        FT x = scale * y + shift;
        result = result + f(x) * w;
    }
    return result * scale;
}
template <class FT, class Q, class F> inline auto integrate_segments(F f, const FT* bnds, int nb, const Q& quad) -> decltype(f(bnds[0])) {
    auto result = integrate1<FT>(f, bnds[0], bnds[1], quad);
    for (int i = 1; i < nb - 1; ++i) result = result + integrate1<FT>(f, bnds[i], bnds[i + 1], quad);
    return result;
}

// ---- terminal velocities                                           P3_terminal_velocity.jl:4-173
template <class FT> struct ChenCurve2 {  // CO.Chen2022VelocityCurve with 2 / 3 terms
    FT a[3], b[3], c[3];
    int n;
    FT operator()(FT D) const {
        FT v = FT(0);
        for (int i = 0; i < n; ++i) v = v + a[i] * pow_(D, b[i]) * exp_(-c[i] * D);
        return v;
    }
};
template <class FT, class P> struct IceVelocity {   // P3IceParticleVelocityFunctor
    ChenCurve2<FT> small_, large_;
    FT cutoff;
    const P3State<FT>* s;
    FT operator()(FT D) const {
        FT v = (D <= cutoff) ? small_(D) : large_(D);
        FT ar = s->prm->aspect_oblate ? FT(cbrt_(phi_i<FT>(*s, D))) : FT(1);
        return v * ar;
    }
};
template <class FT, class P> inline IceVelocity<FT, P> ice_particle_terminal_velocity(const P& p, FT rho_a, const P3State<FT>& s) {
    IceVelocity<FT, P> f;
    const FT rho_i = FT(916.7);   // hard-coded in the reference (P3_terminal_velocity.jl:32)
    chen2022_vel_coeffs_small_ice<FT>(p.vel_small_ice, rho_a, rho_i, f.small_.a, f.small_.b, f.small_.c);
    chen2022_vel_coeffs_large_ice<FT>(p.vel_large_ice, rho_a, rho_i, f.large_.a, f.large_.b, f.large_.c);
    f.small_.n = f.large_.n = 2;
    f.cutoff = FT(p.vel_small_ice.cutoff);
    f.s = &s;
    return f;
}
template <class FT, class P> inline ChenCurve2<FT> rain_particle_terminal_velocity(const P& p, FT rho_a) {
    ChenCurve2<FT> f;
    chen2022_vel_coeffs_rain<FT>(p.vel_rain, rho_a, f.a, f.b, f.c);
    f.n = 3;
    return f;
}
template <class FT, class P>
inline void ice_terminal_velocity_weighted(const P& p, FT rho_a, const P3State<FT>& s, FT logl, FT pq, FT& v_n, FT& v_m) {
    if (s.N_ice < eps_type<FT>() || s.L_ice < eps_type<FT>()) { v_n = v_m = FT(0); return; }
    auto v = ice_particle_terminal_velocity<FT>(p, rho_a, s);
    auto n = ice_psd<FT>(s, logl);
    FT b[5];
    integral_bounds<FT>(s, logl, pq, b);
    v_n = integrate_segments<FT>([&](FT D) { return FT(n(D) * v(D)); }, b, 5, p.quad) / s.N_ice;
    v_m = integrate_segments<FT>([&](FT D) { return FT(n(D) * v(D) * ice_mass<FT>(s, D)); }, b, 5, p.quad) / s.L_ice;
}

// ---- P3.het_ice_nucleation                                                 P3_processes.jl:20-45
template <class FT, class Dd> inline void p3_het_ice_nucleation(const Dd& dust, const Thermo<FT>& tps, FT q_lcl, FT N_lcl, FT RH, FT T, FT rho_a, FT& dNdt, FT& dLdt) {
    FT J = ABIFM_J<FT>(dust, FT(RH - a_w_ice<FT>(tps, T)));
    FT JA = isfinite_(J) ? FT(J * FT(1e-10)) : FT(0);
    dNdt = jmax(FT(0), JA * N_lcl);
    dLdt = jmax(FT(0), JA * q_lcl * rho_a);
}

// ---- P3.ice_melt                                                            P3_processes.jl:64-94
template <class FT, class P> inline void ice_melt(const P& p, FT T, FT rho_a, const P3State<FT>& s, FT logl, FT& dNdt, FT& dLdt) {
    Thermo<FT> tps(p.warm.tps);
    const auto& aps = p.warm.aps;
    FT L_f = tps.L_f(T);
    auto v = ice_particle_terminal_velocity<FT>(p, rho_a, s);
    FT cbrt_Nsc = cbrt_(FT(aps.nu_air) / FT(aps.D_vapor));
    auto F_v = [&](FT D) { return FT(FT(s.prm->vent_a) + FT(s.prm->vent_b) * cbrt_Nsc * sqrt_(D * v(D) / FT(aps.nu_air))); };
    auto n = ice_psd<FT>(s, logl);
    FT fac_ = 4 * FT(aps.K_therm) / L_f * (T - FT(s.prm->T_freeze));
    FT b[5];
    integral_bounds<FT>(s, logl, FT(1e-6), b);
    FT I = integrate_segments<FT>([&](FT D) { return FT(dice_mass_dD<FT>(s, D) * F_v(D) * n(D) / D); }, b, 5, p.quad);
    dLdt = jmax(FT(0), fac_ * I);
    dNdt = s.N_ice / s.L_ice * dLdt;
}

// ---- collisions                                                          P3_processes.jl:112-655
template <class FT> struct Vec3 { FT v[3]; Vec3 operator+(const Vec3& o) const { Vec3 r; for (int i = 0; i < 3; ++i) r.v[i] = v[i] + o.v[i]; return r; }
    Vec3 operator*(FT w) const { Vec3 r; for (int i = 0; i < 3; ++i) r.v[i] = v[i] * w; return r; } };
template <class FT> struct Vec10 { FT v[10]; Vec10 operator+(const Vec10& o) const { Vec10 r; for (int i = 0; i < 10; ++i) r.v[i] = v[i] + o.v[i]; return r; }
    Vec10 operator*(FT w) const { Vec10 r; for (int i = 0; i < 10; ++i) r.v[i] = v[i] * w; return r; } };

// LocalRimeDensity callable                                      CMP/MicrophysicsP3.jl:222-239
template <class FT, class S> inline FT local_rime_density(const S& prm, FT Ri) {
    Ri = jclamp(Ri, FT(1), FT(12));
    auto cl93 = [&](FT R) { return FT(FT(prm.rim_a) + FT(prm.rim_b) * R + FT(prm.rim_c) * (R * R)); };
    if (Ri <= FT(8)) return cl93(Ri);
    FT r8 = cl93(FT(8));
    FT f = (Ri - 8) / (12 - 8);
    return (1 - f) * r8 + f * FT(prm.rim_rho_ice);
}
template <class FT> inline FT volume_sphere_D(FT D) { return D * D * D * pi<FT>() / 6; }

// CM2.pdf_cloud_parameters / size_distribution / get_size_distribution_bounds       CM2:203-355
template <class FT> struct CloudPSD { FT logN0c, lam_c, nu_cD, mu_cD;
    FT operator()(FT D) const { FT v = exp_(logN0c + nu_cD * log_(D) - lam_c * pow_(D, mu_cD)); return (val_(logN0c) == -INFINITY) ? FT(0) : v; } };
template <class FT, class PC> inline CloudPSD<FT> cloud_psd(const PC& pdf_c, FT q, FT rho, FT N) {
    FT lA, lB;
    log_pdf_cloud_parameters_mass<FT>(pdf_c, q, rho, N, lA, lB);
    FT k_m = FT(pdf_c.rho_w) * pi<FT>() / 6;
    CloudPSD<FT> f;
    f.logN0c = lA + log_(FT(3)) + (FT(pdf_c.nu_c) + 1) * log_(k_m);
    f.lam_c = exp_(lB) * pow_(k_m, FT(pdf_c.mu_c));
    f.nu_cD = 3 * FT(pdf_c.nu_c) + 2;
    f.mu_cD = 3 * FT(pdf_c.mu_c);
    return f;
}
template <class FT> struct RainPSD { FT N0r, Dr_mean; FT operator()(FT D) const { FT v = N0r * exp_(-D / Dr_mean); return (N0r == FT(0)) ? FT(0) : v; } };

template <class FT, class P> struct CollisionCtx {
    const P* p;
    const P3State<FT>* s;
    IceVelocity<FT, P> v_ice;
    ChenCurve2<FT> v_liq;
    FT rho_a, T, rho_w;
    // ∂ₜV(Dᵢ, Dₗ) = E K |v_ice - v_liq|, K = evalpoly(Dₗ, (π r², π r, π/4))          :112-135
    FT dV(FT Di, FT Dl) const {
        FT r = sqrt_(ice_area<FT>(*s, Di) / pi<FT>());
        FT K = pi<FT>() * (r * r) + Dl * (pi<FT>() * r + Dl * FT(3.141592653589793238462643383279502884L / 4));
        return K * fabs_(v_ice(Di) - v_liq(Dl));
    }
    // ρ′_rim(Dᵢ, Dₗ)                                                              :152-166
    FT rho_rim_local(FT Di, FT Dl) const {
        FT TC = T - FT(s->prm->T_freeze);
        FT vt = fabs_(v_ice(Di) - v_liq(Dl));
        FT Ri = (Dl * 1000000 * vt) / (2 * TC);
        return local_rime_density<FT>(*s->prm, Ri);
    }
    FT m_liq(FT D) const { return rho_w * volume_sphere_D<FT>(D); }
};

// compute_max_freeze_rate                                                 P3_processes.jl:184-219
template <class FT, class P> inline FT max_freeze_rate(const P& p, const CollisionCtx<FT, P>& c, FT Di) {
    Thermo<FT> tps(p.warm.tps);
    const auto& aps = p.warm.aps;
    FT T = c.T, rho_a = c.rho_a;
    FT T_frz = tps.T_freeze();
    FT Lv = tps.L_v(T), L_f = tps.L_f(T);
    FT dT = T_frz - T;
    FT drho = rho_a * (tps.p2q(T_frz, rho_a, tps.p_sat_ice(T_frz)) - tps.p2q(T, rho_a, tps.p_sat_ice(T)));
    FT denom = L_f - FT(p.warm.tps.cp_l) * dT;
    if (T >= T_frz) return FT(0);
    if (!(denom > FT(0))) return floatmax_<FT>();
    FT cbrt_Nsc = cbrt_(FT(aps.nu_air) / FT(aps.D_vapor));
    FT F_v = FT(c.s->prm->vent_a) + FT(c.s->prm->vent_b) * cbrt_Nsc * sqrt_(Di * c.v_ice(Di) / FT(aps.nu_air));
    return 2 * (pi<FT>() * Di) * F_v * (FT(aps.K_therm) * dT + Lv * FT(aps.D_vapor) * drho) / denom;
}

// closed_rain_inner_NM                                                    P3_processes.jl:344-369
template <class FT, class P>
inline void closed_rain_inner_NM(const CollisionCtx<FT, P>& c, FT v_i, FT r_i, FT D_min, FT D_max, FT N0r, FT Dr_mean, FT& dN, FT& dM) {
    FT lam = 1 / Dr_mean;
    const int maxit = is_float32<FT>() ? 8 : 10;
    FT Dstar = brent_fixed<FT>([&](FT D) { return FT(c.v_liq(D) - v_i); }, D_min, D_max, maxit);   // crossover_diameter :326-335
    FT coef[3] = {pi<FT>() * (r_i * r_i), pi<FT>() * r_i, FT(3.141592653589793238462643383279502884L / 4)};
    auto Ip = [&](FT a, FT b, FT p, FT alpha) {
        FT acc = coef[0] * gamma_inc_moment<FT>(a, b, p, alpha);
        for (int i = 1; i < 3; ++i) acc = acc + coef[i] * gamma_inc_moment<FT>(a, b, p + FT(i), alpha);
        return acc;
    };
    auto flux = [&](FT a, FT b, FT p) {
        FT s = v_i * Ip(a, b, p, lam);
        for (int j = 0; j < 3; ++j) s = s - c.v_liq.a[j] * Ip(a, b, p + c.v_liq.b[j], lam + c.v_liq.c[j]);
        return s;
    };
    auto crossing = [&](FT p) { return FT(flux(D_min, Dstar, p) - flux(Dstar, D_max, p)); };
    FT mfac = c.rho_w * volume_sphere_D<FT>(FT(1));
    dN = N0r * crossing(FT(0));
    dM = N0r * mfac * crossing(FT(3));
}

// ∫liquid_ice_collisions                                                   P3_processes.jl:449-567
template <class FT, class P>
inline Vec10<FT> liquid_ice_collisions(const P& p, const P3State<FT>& s, FT logl, FT L_c, FT N_c, FT L_r, FT N_r, FT rho_a, FT T) {
    CollisionCtx<FT, P> c;
    c.p = &p; c.s = &s; c.rho_a = rho_a; c.T = T; c.rho_w = FT(p.warm.sb.pdf_c.rho_w);
    c.v_ice = ice_particle_terminal_velocity<FT>(p, rho_a, s);
    c.v_liq = rain_particle_terminal_velocity<FT>(p, rho_a);
    auto n_c = cloud_psd<FT>(p.warm.sb.pdf_c, FT(L_c / rho_a), rho_a, N_c);
    RainPDF<FT> rp = pdf_rain_parameters<FT>(p.warm.sb.pdf_r, FT(L_r / rho_a), rho_a, N_r);
    RainPSD<FT> n_r{rp.N0r, rp.Dr_mean};
    auto n_i = ice_psd<FT>(s, logl);
    const FT pq = FT(0.00001);
    FT ib[5];
    integral_bounds<FT>(s, logl, pq, ib);
    FT bc[2] = {generalized_gamma_quantile<FT>(n_c.nu_cD, n_c.mu_cD, n_c.lam_c, pq), generalized_gamma_quantile<FT>(n_c.nu_cD, n_c.mu_cD, n_c.lam_c, FT(1 - pq))};
    FT br[2] = {FT(0), FT(0)};
    if (!(rp.Dr_mean == FT(0))) { br[0] = exponential_quantile<FT>(rp.Dr_mean, pq); br[1] = exponential_quantile<FT>(rp.Dr_mean, FT(1 - pq)); }
    auto cloud_integrals = [&](FT Di) {
        return integrate1<FT>([&](FT D) {
            FT t1 = c.dV(Di, D) * n_c(D);
            FT t2 = t1 * c.m_liq(D);
            FT t3 = t2 / c.rho_rim_local(Di, D);
            return Vec3<FT>{{t1, t2, t3}};
        }, bc[0], bc[1], p.quad);
    };
    auto rain_integrals = [&](FT Di) {   // get_liquid_integrals_rain_closed  :381-415
        Vec3<FT> z{{FT(0), FT(0), FT(0)}};
        if (rp.N0r == FT(0) || !(br[1] > br[0])) return z;
        FT v_i = c.v_ice(Di);
        FT r_i = sqrt_(ice_area<FT>(s, Di) / pi<FT>());
        FT dN, dM;
        closed_rain_inner_NM<FT>(c, v_i, r_i, br[0], br[1], rp.N0r, rp.Dr_mean, dN, dM);
        if (!(isfinite_(dN) && isfinite_(dM))) return z;
        FT dB = integrate1<FT>([&](FT D) { return FT(c.dV(Di, D) * n_r(D) * c.m_liq(D) / c.rho_rim_local(Di, D)); }, br[0], br[1], p.quad);
        return Vec3<FT>{{dN, dM, dB}};
    };
    return integrate_segments<FT>([&](FT Di) {
        Vec3<FT> cc = cloud_integrals(Di), rr = rain_integrals(Di);
        FT M_col = cc.v[1] + rr.v[1];
        FT M_frz = jmin(M_col, max_freeze_rate<FT>(p, c, Di));
        FT f_frz = (M_col == FT(0)) ? FT(0) : FT(M_frz / M_col);
        FT wet = (M_col > M_frz) ? FT(1) : FT(0);
        FT n = n_i(Di);
        return Vec10<FT>{{n * cc.v[1] * f_frz, n * cc.v[1] * (1 - f_frz), n * cc.v[0], n * rr.v[1] * f_frz, n * rr.v[1] * (1 - f_frz),
                          n * rr.v[0], n * M_col, n * cc.v[2] * f_frz, n * rr.v[2] * f_frz, n * wet * M_col}};
    }, ib, 5, p.quad);
}

// bulk_liquid_ice_collision_sources                                       P3_processes.jl:606-655
template <class FT> struct CollisionSources { FT dq_c, dq_r, dN_c, dN_r, dL_rim, dL_ice, dB_rim; };
template <class FT, class P>
inline CollisionSources<FT> bulk_liquid_ice_collision_sources(const P& p, const P3State<FT>& s, FT logl, FT L_c, FT N_c, FT L_r, FT N_r, FT rho_a, FT T) {
    const FT D_shd = FT(1e-3);
    FT rho_w = FT(p.warm.sb.pdf_c.rho_w);
    Vec10<FT> r = liquid_ice_collisions<FT>(p, s, logl, L_c, N_c, L_r, N_r, rho_a, T);
    FT QCFRZ = r.v[0], QCSHD = r.v[1], NCCOL = r.v[2], QRFRZ = r.v[3], QRSHD = r.v[4], NRCOL = r.v[5], M_col = r.v[6], BCCOL = r.v[7], BRCOL = r.v[8], wet = r.v[9];
    FT f_wet = (M_col == FT(0)) ? FT(0) : FT(wet / M_col);
    FT NRSHD = QRSHD / (rho_w * volume_sphere_D<FT>(D_shd));
    FT B_rim = (s.rho_rim == FT(0)) ? FT(0) : FT((s.L_ice * s.F_rim) / s.rho_rim);
    FT QIWET = f_wet * s.L_ice * (1 - s.F_rim) / FT(s.prm->tau_wet);
    FT BIWET = f_wet * (s.L_ice / FT(s.prm->rho_i) - B_rim) / FT(s.prm->tau_wet);
    CollisionSources<FT> o;
    o.dq_c = (-QCFRZ - QCSHD) / rho_a;
    o.dq_r = (-QRFRZ + QCSHD) / rho_a;
    o.dN_c = -NCCOL;
    o.dN_r = -NRCOL + NRSHD;
    o.dL_rim = QCFRZ + QRFRZ + QIWET;
    o.dL_ice = QCFRZ + QRFRZ;
    o.dB_rim = BCCOL + BRCOL + BIWET;
    return o;
}

// ice_self_collection                                                     P3_processes.jl:676-712
template <class FT, class P> inline FT ice_self_collection(const P& p, const P3State<FT>& s, FT logl, FT rho_a) {
    auto n_i = ice_psd<FT>(s, logl);
    auto v = ice_particle_terminal_velocity<FT>(p, rho_a, s);
    FT ib[5];
    integral_bounds<FT>(s, logl, eps_type<FT>(), ib);
    FT total = integrate_segments<FT>([&](FT D1) {
        FT v1 = v(D1);
        FT r1 = sqrt_(ice_area<FT>(s, D1) / pi<FT>());
        auto integrand = [&](FT D2) {
            FT v2 = v(D2);
            FT r2 = sqrt_(ice_area<FT>(s, D2) / pi<FT>());
            FT K = pi<FT>() * ((r1 + r2) * (r1 + r2));
            return FT(K * fabs_(v1 - v2) * n_i(D2));
        };
        FT rate = integrate1<FT>(integrand, ib[0], D1, p.quad) + integrate1<FT>(integrand, D1, ib[4], p.quad);
        return FT(rate * n_i(D1));
    }, ib, 5, p.quad);
    return FT(0.5) * total;
}

// ---- IN.liquid_freezing_rate (rain / cloud PSD), immersion_limit_rate, deposition_rate   IN:274-511
template <class FT, class P> inline void rain_freezing_rate(const P& p, FT q, FT rho, FT N, FT T, FT& dn, FT& dq) {
    const FT e = eps_type<FT>();
    FT T_freeze = FT(p.warm.tps.T_freeze);
    FT rho_w = FT(p.warm.sb.pdf_r.rho_w);
    FT n = N / rho;
    RainPDF<FT> rp = pdf_rain_parameters<FT>(p.warm.sb.pdf_r, q, rho, N);
    FT J = FT(p.rain_freezing_het_B) * exp_(FT(p.rain_freezing_het_a) * (T_freeze - T));
    FT M3 = exponential_Mn<FT>(rp.Dr_mean, n, 3), M6 = exponential_Mn<FT>(rp.Dr_mean, n, 6);
    FT V1 = pi<FT>() / 6;
    FT a = J * V1 * M3, b = J * rho_w * (V1 * V1) * M6;
    bool cond = (n > e) && (q > e) && (T < T_freeze - 4);
    dn = cond ? a : FT(0);
    dq = cond ? b : FT(0);
}
template <class FT, class P> inline void cloud_freezing_rate(const P& p, FT q, FT rho, FT N, FT T, FT& dn, FT& dq) {
    const FT e = eps_type<FT>();
    FT T_freeze = FT(p.warm.tps.T_freeze);
    FT rho_w = FT(p.warm.sb.pdf_c.rho_w);
    FT n = N / rho;
    auto c = cloud_psd<FT>(p.warm.sb.pdf_c, q, rho, N);
    FT J = FT(p.rain_freezing_het_B) * exp_(FT(p.rain_freezing_het_a) * (T_freeze - T));
    FT M3 = generalized_gamma_Mn<FT>(c.nu_cD, c.mu_cD, c.lam_c, n, FT(3));
    FT M6 = generalized_gamma_Mn<FT>(c.nu_cD, c.mu_cD, c.lam_c, n, FT(6));
    FT V1 = pi<FT>() / 6;
    FT a = J * V1 * M3, b = J * rho_w * (V1 * V1) * M6;
    bool cond = (n > e) && (q > e) && (T < T_freeze - 4);
    dn = cond ? a : FT(0);
    dq = cond ? b : FT(0);
}
template <class FT, class F> inline FT immersion_limit_rate(const F& opt, FT T, FT rho, FT tau, FT shift, FT n_active) {
    if (T >= FT(opt.T_freeze)) return FT(0);
    FT inpc = exp_(INP_concentration_mean<FT>(opt, T) + shift) / rho;
    return jmax(FT(0), inpc - n_active) / tau;
}
template <class FT, class P>
inline void f23_deposition_rate(const P& p, FT T, FT rho, FT q_tot, FT q_liq, FT q_ice, FT n_ice, FT m_nuc, FT tau_act, FT shift, FT& dn, FT& dq) {
    Thermo<FT> tps(p.warm.tps);
    const auto& opt = p.ice_nucleation;
    FT T_thresh = FT(opt.T_freeze) - 15, S_thresh = FT(0.05);
    FT q_sat_ice = tps.q_sat_ice(T, rho);
    FT q_vap = Thermo<FT>::q_vap(q_tot, q_liq, q_ice);
    FT S_i = q_vap / q_sat_ice - 1;
    bool cond = (T < T_thresh) && (S_i > S_thresh);
    FT inpc = exp_(INP_concentration_mean<FT>(opt, T) + shift) / rho;
    FT a = jmax(FT(0), inpc - n_ice) / tau_act;
    dn = cond ? a : FT(0);
    FT q_excess = jmax(FT(0), q_vap - q_sat_ice);
    dq = jmin(m_nuc * dn, q_excess / (2 * tau_act));
}

// ---- BMT.bulk_microphysics_tendencies(::Microphysics2Moment, mp{WR,<:P3IceParams}, ...)   BMT:898-1083
template <class FT> struct BMT2MP3Out { FT dq_lcl, dn_lcl, dq_rai, dn_rai, dq_ice, dn_ice, dq_rim, db_rim, dn_act; };
template <class FT, class P>
inline BMT2MP3Out<FT> bmt2m_p3(const P& p, FT rho, FT T, FT q_tot, FT q_lcl, FT n_lcl, FT q_rai, FT n_rai, FT q_ice, FT n_ice, FT q_rim, FT b_rim,
                               FT logl, FT shift) {
    const FT eM = eps_type<FT>(), eN = eps_type<FT>();
    rho = clamp_to_nonneg(rho); q_tot = clamp_to_nonneg(q_tot); q_lcl = clamp_to_nonneg(q_lcl); q_rai = clamp_to_nonneg(q_rai);
    n_lcl = clamp_to_nonneg(n_lcl); n_rai = clamp_to_nonneg(n_rai); q_ice = clamp_to_nonneg(q_ice); n_ice = clamp_to_nonneg(n_ice);
    q_rim = clamp_to_nonneg(q_rim); b_rim = clamp_to_nonneg(b_rim);
    FT L_lcl = q_lcl * rho, L_rai = q_rai * rho, N_lcl = n_lcl * rho, N_rai = n_rai * rho, L_ice = q_ice * rho, N_ice = n_ice * rho,
       L_rim = q_rim * rho, B_rim = b_rim * rho;
    P3State<FT> s = state_from_prognostic<FT>(p.scheme, L_ice, N_ice, L_rim, B_rim);
    Thermo<FT> tps(p.warm.tps);
    BMT2MP3Out<FT> o;
    Warm2MOut<FT> w = warm_rain_tendencies_2m<FT>(p.warm, T, q_tot, q_lcl, q_rai, q_ice, rho, n_lcl, n_rai);
    o.dq_lcl = w.dq_lcl_dt; o.dn_lcl = w.dn_lcl_dt; o.dq_rai = w.dq_rai_dt; o.dn_rai = w.dn_rai_dt;
    o.dq_ice = o.dn_ice = o.dq_rim = o.db_rim = FT(0);
    o.dn_act = FT(0);
    if (q_ice > eM && n_ice > eN) {
        CollisionSources<FT> c = bulk_liquid_ice_collision_sources<FT>(p, s, logl, L_lcl, N_lcl, L_rai, N_rai, rho, T);
        o.dq_lcl += c.dq_c; o.dq_rai += c.dq_r; o.dn_lcl += c.dN_c / rho; o.dn_rai += c.dN_r / rho;
        o.dq_ice += c.dL_ice / rho; o.dq_rim += c.dL_rim / rho; o.db_rim += c.dB_rim / rho;
        FT agg = ice_self_collection<FT>(p, s, logl, rho);
        o.dn_ice -= agg / rho;
        FT mN = FT(0), mL = FT(0);
        if (T > tps.T_freeze()) ice_melt<FT>(p, T, rho, s, logl, mN, mL);
        FT dq_m = mL / rho, dn_m = mN / rho;
        o.dq_rai += dq_m; o.dn_rai += dn_m; o.dq_ice -= dq_m; o.dn_ice -= dn_m;
        o.dq_rim -= dq_m * s.F_rim;
        o.db_rim -= (s.rho_rim > FT(0)) ? FT(dq_m * s.F_rim / s.rho_rim) : FT(0);
    }
    FT tau_act = FT(p.tau_act);
    FT m_nuc = FT(p.scheme.rho_i) * volume_sphere_D<FT>(FT(10e-6));
    FT dn_dep, dq_dep;
    f23_deposition_rate<FT>(p, T, rho, q_tot, q_lcl + q_rai, q_ice, n_ice, m_nuc, tau_act, shift, dn_dep, dq_dep);
    o.dn_ice += dn_dep; o.dq_ice += dq_dep;
    FT bn, bq;
    cloud_freezing_rate<FT>(p, q_lcl, rho, N_lcl, T, bn, bq);
    FT cap = immersion_limit_rate<FT>(p.ice_nucleation, T, rho, tau_act, shift, n_ice);
    FT dn_imm = jmin(bn, cap);
    FT dq_imm = (bn > FT(0)) ? FT(bq * dn_imm / bn) : FT(0);
    o.dq_lcl -= dq_imm; o.dn_lcl -= dn_imm; o.dq_ice += dq_imm; o.dn_ice += dn_imm; o.dq_rim += dq_imm; o.db_rim += dq_imm / FT(p.scheme.rho_i);
    FT n_per_q = (q_ice > eM) ? FT(n_ice / q_ice) : FT(0);
    FT dq_d = conv_q_vap_to_q_icl_const<FT>(FT(p.warm.subdep_tau_relax), tps, q_tot, q_lcl, q_ice, q_rai, FT(0), rho, T);
    dq_d = (T > tps.T_freeze()) ? jmin(dq_d, FT(0)) : dq_d;
    FT dn_d = (dq_d < FT(0)) ? FT(n_per_q * dq_d) : FT(0);
    o.dq_ice += dq_d; o.dn_ice += dn_d;
    FT dq_sub = jmin(dq_d, FT(0));
    o.dq_rim += dq_sub * s.F_rim;
    o.db_rim += (s.rho_rim > FT(0)) ? FT(dq_sub * s.F_rim / s.rho_rim) : FT(0);
    o.dn_ice += number_tendency_from_mass_limits<FT>(FT(1e-12), FT(1e-5), FT(100), q_ice, n_ice);
    FT rn, rq;
    rain_freezing_rate<FT>(p, q_rai, rho, N_rai, T, rn, rq);
    o.dq_rai -= rq; o.dn_rai -= rn; o.dq_ice += rq; o.dn_ice += rn; o.dq_rim += rq; o.db_rim += rq / FT(p.scheme.rho_i);
    return o;
}

}  // namespace orc
