// kernels_icenuc.cu — ice-nucleation rates, water activities, ARG2000 aerosol activation and
// the fused "ice nucleation + ARG" kernel of BASELINE config 3; C-ABI in include/cumicro.h.
#include <cmath>
#include <limits>

#include "cm_icenuc.cuh"
#include "cm_launch.cuh"

// config 3 sweeps: grid-stride 128x4 3.50 ms, x6 3.09, x8 3.10 (round 1); tile shape 128x5 1.945, x6 1.840, x7 1.894; with the
// final body 128x6 1.547, 384x2 1.537, 768x1 1.542, 896x1 1.513 (one block per SM: the 58 KB body overflows the instruction cache)
#ifndef CUMICRO_ARG_MINB
#define CUMICRO_ARG_MINB 1
#endif
#ifndef CUMICRO_ARG_BLOCK
#define CUMICRO_ARG_BLOCK 896
#endif
#ifndef CUMICRO_ARG_TILED
#define CUMICRO_ARG_TILED 1   /* config 3: 1.983 (grid-stride register-loading shape) -> 1.840 ms (bulk-copied tiles, cm_launch.cuh) */
#endif
namespace {

using namespace cm;
using D = double;
template <class FT> constexpr bool is_f32() { return sizeof(FT) == 4; }
template <class FT> struct PI;
template <> struct PI<double> { using params = cumicro_params_icenuc_f64; };
template <> struct PI<float> { using params = cumicro_params_icenuc_f32; };

struct IceNucBase {
    cumicro_params_icenuc_f64 p;
    ThermoK<D> tk;
    ArgK<D> k;
    unsigned long long* n_domain_errors;   // device counter (may be NULL)
    __device__ __forceinline__ void flag() const {
        if (n_domain_errors) atomicAdd(n_domain_errors, 1ULL);
    }
};

// pointwise leaf functions: out = fn(x [, y]);  `what` as in include/cumicro.h
struct IceNucLeaf : IceNucBase {
    int what;
    __device__ __forceinline__ void operator()(const D (&x)[2], D (&y)[1]) const {
        const D nan = __longlong_as_double(0x7ff8000000000000LL);
        D v = 0;
        bool err = false;
        switch (what) {
            case 0: v = deposition_J<D>(p.dust, x[0], k.ln10); break;
            case 1: v = ABIFM_J<D>(p.dust, x[0], k.ln10); break;
            case 2: v = homogeneous_J_cubic<D>(p.koop, x[0], k.ln10, err); break;
            case 3: v = homogeneous_J_linear<D>(p.koop, x[0], k.ln10); break;
            case 4: case 5: case 6: case 7: {
                const TempState<D> ts = temp_state(tk, x[0]);
                const D pl = p_sat_liq(tk, ts);
                if (what == 4) v = p_sat_ice(tk, ts) / pl;                       // CO.a_w_ice
                else if (what == 5) v = x[1] / pl;                              // CO.a_w_eT(e, T)
                else {
                    const auto& h = p.h2so4;                                     // CO.H2SO4_soln_saturation_vapor_pressure
                    const D xf = x[1], wh = h.w_2 * xf;
                    const D psol = exp_full_(h.c[0] - h.c[1] * xf + h.c[2] * xf * wh - h.c[3] * xf * (wh * wh) +
                                             (h.c[4] + h.c[5] * xf - h.c[6] * xf * wh) / x[0]) * 100.0;
                    v = (what == 6) ? psol / pl : psol;                         // CO.a_w_xT / p_sol
                }
                break;
            }
            case 8: {                                                            // IN.P3_deposition_N_i
                const auto& ip = p.mm2014;
                const D Tp = fmax_(ip.T_dep_thres, x[0]);
                const D Ni = 1000.0 * ip.c1 * exp_full_(ip.c2 * (ip.T0 - Tp));
                v = (x[0] < ip.T0) ? Ni : 0.0;
                break;
            }
            case 9: {                                                            // IN.INP_concentration_mean
                const D Tc = fmin_(x[0] - p.frostenberg.T_freeze, 0.0);
                v = 9.0 * log_full_(-p.frostenberg.b * Tc / 10.0) - p.frostenberg.log_a;
                break;
            }
            case 10: {                                                           // IN.dust_activated_number_fraction(Si, T)
                err = !(x[0] < p.mohler.Si_max);
                const bool warm = x[1] > p.mohler.T_thr;
                const D S0 = warm ? p.dust.S0_warm : p.dust.S0_cold;
                const D a = warm ? p.dust.a_warm : p.dust.a_cold;
                v = clamp0_(exp_full_(a * (x[0] - S0)) - 1.0);
                break;
            }
            default: break;
        }
        if (err) { v = nan; flag(); }
        y[0] = v;
    }
};

// multi-argument nucleation rates: (in[0..4]) -> (out, out2);  `what` as in include/cumicro.h (cumicro_icenuc_rates_*)
struct IceNucRates : IceNucBase {
    int what;
    __device__ __forceinline__ void operator()(const D (&x)[5], D (&y)[2]) const {
        const D nan = __longlong_as_double(0x7ff8000000000000LL);
        D v = 0, v2 = 0;
        bool err = false;
        switch (what) {
            case 0: {                                                            // IN.MohlerDepositionRate(Si, T, dSi_dt, N_aer)
                err = !(x[0] < p.mohler.Si_max);
                const D a = (x[1] > p.mohler.T_thr) ? p.dust.a_warm : p.dust.a_cold;
                v = clamp0_(x[3] * a * x[2]);
                break;
            }
            case 1: {                                                            // IN.P3_het_N_i(T, N_l, V_l, dt)
                const auto& ip = p.mm2014;
                v = x[1] * (1.0 - exp_full_(-ip.het_B * x[2] * x[3] * exp_full_(ip.het_a * (ip.T0 - x[0]))));
                break;
            }
            case 2: {                                                            // IN.INP_concentration_frequency(INPC, T)
                const auto& f = p.frostenberg;
                const D Tc = fmin_(x[1] - f.T_freeze, 0.0);
                const D mu = 9.0 * log_full_(-f.b * Tc / 10.0) - f.log_a;
                const D s2 = f.sigma * f.sigma, d = log_full_(x[0]) - mu;
                v = (x[1] >= f.T_freeze) ? 0.0 : exp_full_(-(d * d) / (2.0 * s2)) / sqrt_(3.141592653589793 * 2.0 * s2);
                break;
            }
            case 3: {                                                            // P3.het_ice_nucleation(q_lcl, N_lcl, RH, T, rho)
                const TempState<D> ts = temp_state(tk, x[3]);
                const D a_w_ice = p_sat_ice(tk, ts) / p_sat_liq(tk, ts);
                const D J = ABIFM_J<D>(p.dust, x[2] - a_w_ice, k.ln10);
                const D JA = isfinite(J) ? J * 1e-10 : 0.0;
                v = clamp0_(JA * x[1]);
                v2 = clamp0_(JA * x[0] * x[4]);
                break;
            }
            default: break;
        }
        if (err) { v = nan; flag(); }
        y[0] = v;
        y[1] = v2;
    }
};

// ARG2000 + nucleation rates, MODES aerosol modes:
//   in : T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice
//   out: S_max, N_act[MODES], M_act[MODES], J_dep, J_ABIFM, J_hom, Δa_w   (NULL columns skipped)
template <int MODES, bool WANT_M> struct ArgIceNuc : IceNucBase {
    __device__ __forceinline__ void operator()(const D (&x)[8], D (&y)[1 + 2 * MODES + 4]) const {
        const ArgOut o = arg2000<WANT_M, true, (MODES <= 4 ? MODES : -1)>(p, tk, k, x[0],   // MODES = 8 serves n_modes 5..8: run-time bound
             x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
        y[0] = o.S_max;
#pragma unroll
        for (int i = 0; i < MODES; ++i) {
            y[1 + i] = o.N_act[i];
            y[1 + MODES + i] = WANT_M ? o.M_act[i] : 0.0;
        }
        bool err = false;
        y[1 + 2 * MODES + 0] = deposition_J<D>(p.dust, o.da_w, k.ln10);
        y[1 + 2 * MODES + 1] = ABIFM_J<D>(p.dust, o.da_w, k.ln10);
        D jh;
        if (p.hom_linear) {
            jh = homogeneous_J_linear<D>(p.koop, o.da_w, k.ln10);
        } else {
            jh = homogeneous_J_cubic<D>(p.koop, o.da_w, k.ln10, err);
            if (err) { jh = __longlong_as_double(0x7ff8000000000000LL); flag(); }   // the reference throws DomainError (IN:558-562)
        }
        y[1 + 2 * MODES + 2] = jh;
        y[1 + 2 * MODES + 3] = o.da_w;
    }
};

template <class FT, class F> F make_icenuc(const typename PI<FT>::params* p, unsigned long long* counter) {
    F f{};
    widen(*p, f.p);
    f.tk = make_thermo_k<D>(f.p.tps, is_f32<FT>());
    f.k = make_arg_k<D>(f.p, is_f32<FT>());
    f.n_domain_errors = counter;
    return f;
}

template <class FT>
int icenuc_leaf_impl(const typename PI<FT>::params* p, int what, int64_t n, const FT* x, const FT* y, FT* out,
                     unsigned long long* n_domain_errors, void* stream) {
    if (what < 0 || what > 10) return cmh::fail(CUMICRO_E_OPTION, "icenuc: what = %d (expected 0..10)", what);
    const bool two = (what == 5 || what == 6 || what == 7 || what == 10);
    const FT* in[2] = {x, two ? y : x};
    FT* o[1] = {out};
    int st = validate_columns<FT, 2>(p, n, in);
    if (st) return st;
    if ((st = require_outputs<FT, 1>(n, o, 1))) return st;
    IceNucLeaf f = make_icenuc<FT, IceNucLeaf>(p, n_domain_errors);
    f.what = what;
    return launch_pointwise<FT, 2, 1, IceNucLeaf, 256, 2>(f, n, in, o, (cudaStream_t)stream, "icenuc leaf launch");
}

template <class FT>
int icenuc_rates_impl(const typename PI<FT>::params* p, int what, int64_t n, const FT* const* in5, FT* out, FT* out2,
                      unsigned long long* n_domain_errors, void* stream) {
    static const int nin[4] = {4, 4, 2, 5};
    if (what < 0 || what > 3) return cmh::fail(CUMICRO_E_OPTION, "icenuc_rates: what = %d (expected 0..3)", what);
    if (in5 == nullptr) return cmh::fail(CUMICRO_E_NULL, "icenuc_rates: column pointer table is NULL");
    const FT* in[5];
    for (int c = 0; c < 5; ++c) in[c] = (c < nin[what]) ? in5[c] : in5[0];
    FT* o[2] = {out, out2};
    int st = validate_columns<FT, 5>(p, n, in);
    if (st) return st;
    if ((st = require_outputs<FT, 2>(n, o, 1))) return st;
    IceNucRates f = make_icenuc<FT, IceNucRates>(p, n_domain_errors);
    f.what = what;
    return launch_pointwise<FT, 5, 2, IceNucRates, 256, 2>(f, n, in, o, (cudaStream_t)stream, "icenuc rates launch");
}

template <class FT, int MODES>
int arg_icenuc_launch(const typename PI<FT>::params* p, int64_t n, const FT* const (&in)[8], FT* S_max, FT* const* N_act,
                      FT* const* M_act, FT* J_dep, FT* J_abifm, FT* J_hom, FT* da_w, unsigned long long* counter, cudaStream_t s) {
    FT* out[1 + 2 * MODES + 4];
    out[0] = S_max;
    bool want_m = false;
    for (int i = 0; i < MODES; ++i) {
        out[1 + i] = N_act ? N_act[i] : nullptr;
        out[1 + MODES + i] = M_act ? M_act[i] : nullptr;
        want_m = want_m || (out[1 + MODES + i] != nullptr);
    }
    out[1 + 2 * MODES + 0] = J_dep;
    out[1 + 2 * MODES + 1] = J_abifm;
    out[1 + 2 * MODES + 2] = J_hom;
    out[1 + 2 * MODES + 3] = da_w;
#if CUMICRO_ARG_TILED
    if (want_m)
        return launch_pointwise_tiled<FT, 8, 1 + 2 * MODES + 4, ArgIceNuc<MODES, true>, CUMICRO_ARG_BLOCK, CUMICRO_ARG_MINB>(
            make_icenuc<FT, ArgIceNuc<MODES, true>>(p, counter), n, in, out, s, "arg_icenuc launch");
    return launch_pointwise_tiled<FT, 8, 1 + 2 * MODES + 4, ArgIceNuc<MODES, false>, CUMICRO_ARG_BLOCK, CUMICRO_ARG_MINB>(
        make_icenuc<FT, ArgIceNuc<MODES, false>>(p, counter), n, in, out, s, "arg_icenuc launch");
#else
    if (want_m)
        return launch_pointwise<FT, 8, 1 + 2 * MODES + 4, ArgIceNuc<MODES, true>, 128, 6, false>(
            make_icenuc<FT, ArgIceNuc<MODES, true>>(p, counter), n, in, out, s, "arg_icenuc launch");
    return launch_pointwise<FT, 8, 1 + 2 * MODES + 4, ArgIceNuc<MODES, false>, 128, 6, false>(
        make_icenuc<FT, ArgIceNuc<MODES, false>>(p, counter), n, in, out, s, "arg_icenuc launch");
#endif
}

template <class FT>
int arg_icenuc_impl(const typename PI<FT>::params* p, int64_t n, const FT* T, const FT* pr, const FT* w, const FT* q_tot,
                    const FT* q_liq, const FT* q_ice, const FT* N_liq, const FT* N_ice, FT* S_max, FT* const* N_act,
                    FT* const* M_act, FT* J_dep, FT* J_abifm, FT* J_hom, FT* da_w, unsigned long long* counter, void* stream) {
    const FT* in[8] = {T, pr, w, q_tot, q_liq, q_ice, N_liq, N_ice};
    int st = validate_columns<FT, 8>(p, n, in);
    if (st) return st;
    if (p->n_modes < 1 || p->n_modes > kMaxModes)
        return cmh::fail(CUMICRO_E_OPTION, "n_modes = %d (expected 1..%d)", (int)p->n_modes, kMaxModes);
    cudaStream_t s = (cudaStream_t)stream;
    switch (p->n_modes) {
        case 1: return arg_icenuc_launch<FT, 1>(p, n, in, S_max, N_act, M_act, J_dep, J_abifm, J_hom, da_w, counter, s);
        case 2: return arg_icenuc_launch<FT, 2>(p, n, in, S_max, N_act, M_act, J_dep, J_abifm, J_hom, da_w, counter, s);
        case 3: return arg_icenuc_launch<FT, 3>(p, n, in, S_max, N_act, M_act, J_dep, J_abifm, J_hom, da_w, counter, s);
        case 4: return arg_icenuc_launch<FT, 4>(p, n, in, S_max, N_act, M_act, J_dep, J_abifm, J_hom, da_w, counter, s);
        default: return arg_icenuc_launch<FT, 8>(p, n, in, S_max, N_act, M_act, J_dep, J_abifm, J_hom, da_w, counter, s);
    }
}

}  // namespace

extern "C" {

#define CUMICRO_DEF_ICENUC(SUF, FT)                                                                                    \
    int cumicro_icenuc_##SUF(const cumicro_params_icenuc_##SUF* p, int what, int64_t n, const FT* x, const FT* y,       \
                             FT* out, unsigned long long* n_domain_errors, void* stream) {                             \
        return icenuc_leaf_impl<FT>(p, what, n, x, y, out, n_domain_errors, stream);                                   \
    }                                                                                                                  \
    int cumicro_arg_icenuc_##SUF(const cumicro_params_icenuc_##SUF* p, int64_t n, const FT* T, const FT* pr,            \
                                 const FT* w, const FT* q_tot, const FT* q_liq, const FT* q_ice, const FT* N_liq,      \
                                 const FT* N_ice, FT* S_max, FT* const* N_act, FT* const* M_act, FT* J_dep,            \
                                 FT* J_abifm, FT* J_hom, FT* da_w, unsigned long long* n_domain_errors, void* stream) {\
        return arg_icenuc_impl<FT>(p, n, T, pr, w, q_tot, q_liq, q_ice, N_liq, N_ice, S_max, N_act, M_act, J_dep,      \
                                   J_abifm, J_hom, da_w, n_domain_errors, stream);                                     \
    }
CUMICRO_DEF_ICENUC(f64, double)
CUMICRO_DEF_ICENUC(f32, float)

int cumicro_icenuc_rates_f64(const cumicro_params_icenuc_f64* p, int what, int64_t n, const double* const* in5, double* out,
                             double* out2, unsigned long long* n_domain_errors, void* stream) {
    return icenuc_rates_impl<double>(p, what, n, in5, out, out2, n_domain_errors, stream);
}
int cumicro_icenuc_rates_f32(const cumicro_params_icenuc_f32* p, int what, int64_t n, const float* const* in5, float* out,
                             float* out2, unsigned long long* n_domain_errors, void* stream) {
    return icenuc_rates_impl<float>(p, what, n, in5, out, out2, n_domain_errors, stream);
}

}  // extern "C"
