// kernels_1m.cu — 1-moment bulk microphysics kernels and their C-ABI entry points
// (include/cumicro.h): Instantaneous / InstantaneousVerbose / LinearizedAverage tendencies
// (BMT:505-632) and the 1-moment / non-equilibrium terminal velocities.
#include <cmath>
#include <limits>

#include "cm_1m.cuh"
#include "cm_hostpipe.cuh"
#include "cm_launch.cuh"
#include "cm_sb2006.cuh"

// Block size sweep at 2^24 points (tools/tune_1m.py; blocks x resident blocks = 1024 threads per SM throughout):
//   128x8: Instantaneous 0.944, Verbose 1.127, LinearizedAverage 1.566 ms | 256x4: 0.954, 1.133, 1.529 | 512x2: 0.967, 1.157, 1.505
//   | 1024x1: 0.989, 1.233, 1.469.  Only the largest body (LinearizedAverage, ~60 KB of code) gains from one block per SM
//   (its warps stay in phase and share instruction-cache lines, cf. kernels_fused.cu).
#ifndef CUMICRO_1M_BLOCK
#define CUMICRO_1M_BLOCK 128
#endif
#ifndef CUMICRO_1ML_BLOCK
#define CUMICRO_1ML_BLOCK 896   /* final body, 2^24 points, nsub 1: 1024x1 (64 registers, 456 B spilled) 1.037 ms, 896x1 (72, 248 B) 1.008, 768x1 (85, 208 B) 1.010, 640x1 (102, 100 B) 1.047 */
#endif
#ifndef CUMICRO_1ML_MINB
#define CUMICRO_1ML_MINB 1
#endif
#ifndef CUMICRO_1ML_PIPE
#define CUMICRO_1ML_PIPE 0   /* inputs of the next grid-stride item fetched by cp.async while the current one is computed: 896x1 1.055 ms, 1024x1 1.149, 768x1 1.009 against 1.008 without (the exposed load latency is not what bounds it) */
#endif
#ifndef CUMICRO_1MV_MINB
#define CUMICRO_1MV_MINB 8   /* verbose: 4 -> 1.75 ms, 6 -> 1.34, 8 -> 1.22; linavg: 3.07, 3.01, 2.94 */
#endif
#ifndef CUMICRO_1M_MINB
#define CUMICRO_1M_MINB 7   /* sweep at 2^24 points: 5 -> 1.224 ms, 6 -> 1.118, 8 -> 1.043; later: 7 -> 0.937, 8 -> 0.916; tile shape: 6 -> 0.839, 7 -> 0.832, 8 -> 0.929 */
#endif
// tile shape (cm_launch.cuh, pointwise_kernel_tiled: bulk-copied input tiles, block-uniform loop) or the grid-stride register-loading shape
#ifndef CUMICRO_1M_TILED
#define CUMICRO_1M_TILED 1    /* Instantaneous 2^24 points: 0.892 (grid-stride, 128x8) -> 0.832 ms (tiles, 128x7) */
#endif
#ifndef CUMICRO_1MV_TILED
#define CUMICRO_1MV_TILED 1   /* InstantaneousVerbose (22 columns out): 1.157 -> 0.902 ms (128x8; 128x7 0.906) */
#endif
#ifndef CUMICRO_1ML_TILED
#define CUMICRO_1ML_TILED 0   /* LinearizedAverage nsub = 1: 1.266 ms as it is (1024x1, grid-stride); tiles: 1024x1 1.457, 512x2 1.400, 128x7 1.320 */
#endif
namespace {

using namespace cm;

// Functors compute in Float64; entry points are templated on the column type FT.
using D = double;
template <class FT> constexpr bool is_f32() { return sizeof(FT) == 4; }
struct OneMBase {
    P<D>::params_1m p;
    ThermoK<D> tk;
    OneMK<D> k;
};

// BMT:505-514 — 7 columns in, 4 tendencies out
// STD: the default exponent structure (cm_1m.cuh, OneMK::std_exponents): powers of λ⁻¹ by multiplication
template <bool STD> struct OneMInst : OneMBase {
    __device__ __forceinline__ void operator()(const D (&x)[7], D (&y)[4]) const {
        const Src1M<D> r = microphysics_source_terms_1m<D, STD>(p, tk, k, x[0], x[1], x[2], x[3], x[4], x[5], x[6]);
        aggregate_tendencies_1m<D>(r, y);
    }
};
// BMT:533-543 — 4 tendencies + the 18 source terms
template <bool STD> struct OneMVerbose : OneMBase {
    __device__ __forceinline__ void operator()(const D (&x)[7], D (&y)[4 + S1M_NSRC]) const {
        const Src1M<D> r = microphysics_source_terms_1m<D, STD>(p, tk, k, x[0], x[1], x[2], x[3], x[4], x[5], x[6]);
        D t[4];
        aggregate_tendencies_1m<D>(r, t);
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = t[i];
#pragma unroll
        for (int i = 0; i < S1M_NSRC; ++i) y[4 + i] = r.s[i];
    }
};
// BMT:572-632 — nsub linearised implicit substeps
template <bool STD> struct OneMLinAvg : OneMBase {
    LinAvgK<D> lk;
    D Lv_over_cp, Ls_over_cp;
    int nsub;
    __device__ __forceinline__ void operator()(const D (&x)[7], D (&y)[4]) const {
        bmt1m_linearized_average<D, STD>(p, tk, k, x[0], x[1], x[2], x[3], x[4], x[5], x[6], lk, nsub, Lv_over_cp, Ls_over_cp, y);
    }
};

template <class FT> int check_1m_options(const typename P<FT>::params_1m* p) {
    const auto& o = p->processes;
    const int32_t v[13] = {o.cloud_liquid_formation, o.cloud_ice_formation, o.cloud_ice_melt, o.rain_autoconversion,
                           o.snow_autoconversion, o.rain_condensation_evaporation, o.snow_deposition_sublimation, o.snow_melt,
                           o.cloud_liquid_rain_accretion, o.cloud_liquid_snow_accretion, o.cloud_ice_rain_accretion,
                           o.cloud_ice_snow_accretion, o.rain_snow_accretion};
    const int32_t hi[13] = {1, 2, 1, 2, 2, 1, 2, 1, 1, 1, 1, 1, 1};
    for (int i = 0; i < 13; ++i)
        if (v[i] < 0 || v[i] > hi[i]) return cmh::fail(CUMICRO_E_OPTION, "processes slot %d = %d (expected 0..%d)", i, (int)v[i], (int)hi[i]);
    return CUMICRO_OK;
}

template <class FT, class F> F make_1m(const typename P<FT>::params_1m* p) {
    F f{};
    widen(*p, f.p);   // exact for Float32 blocks
    f.tk = make_thermo_k<D>(f.p.tps, is_f32<FT>());
    f.k = make_1m_k<D>(f.p, is_f32<FT>());
    return f;
}

// one mode with the body variant STD (default exponent structure or not) decided on the host
template <class FT, bool STD>
int bmt1m_launch(int mode, const typename P<FT>::params_1m* p, int64_t n, const FT* const (&in)[7], FT dt, int nsub, FT* const (&o4)[4],
                 FT* const* src18, cudaStream_t s) {
    if (mode == 0) {
        using F = OneMInst<STD>;
#if CUMICRO_1M_TILED
        return launch_pointwise_tiled<FT, 7, 4, F, CUMICRO_1M_BLOCK, CUMICRO_1M_MINB>(make_1m<FT, F>(p), n, in, o4, s, "bmt1m_inst launch");
#else
        return launch_pointwise<FT, 7, 4, F, CUMICRO_1M_BLOCK, CUMICRO_1M_MINB, false>(make_1m<FT, F>(p), n, in, o4, s, "bmt1m_inst launch");
#endif
    } else if (mode == 1) {
        using F = OneMVerbose<STD>;
        FT* o22[4 + S1M_NSRC];
        for (int i = 0; i < 4; ++i) o22[i] = o4[i];
        for (int i = 0; i < S1M_NSRC; ++i) o22[4 + i] = src18[i];
#if CUMICRO_1MV_TILED
        return launch_pointwise_tiled<FT, 7, 4 + S1M_NSRC, F, CUMICRO_1M_BLOCK, CUMICRO_1MV_MINB>(make_1m<FT, F>(p), n, in, o22, s, "bmt1m_verbose launch");
#else
        return launch_pointwise<FT, 7, 4 + S1M_NSRC, F, CUMICRO_1M_BLOCK, CUMICRO_1MV_MINB, false>(make_1m<FT, F>(p), n, in, o22, s, "bmt1m_verbose launch");
#endif
    } else {
        using F = OneMLinAvg<STD>;
        F f = make_1m<FT, F>(p);
        f.lk = make_linavg_k<D>((D)dt, nsub);
        f.nsub = nsub;
        f.Lv_over_cp = f.p.tps.LH_v0 / f.p.tps.cp_d;
        f.Ls_over_cp = f.p.tps.LH_s0 / f.p.tps.cp_d;
#if CUMICRO_1ML_TILED
        return launch_pointwise_tiled<FT, 7, 4, F, CUMICRO_1ML_BLOCK, CUMICRO_1ML_MINB>(f, n, in, o4, s, "bmt1m_linavg launch");
#else
        return launch_pointwise<FT, 7, 4, F, CUMICRO_1ML_BLOCK, CUMICRO_1ML_MINB, false, CUMICRO_1ML_PIPE != 0>(f, n, in, o4, s, "bmt1m_linavg launch");
#endif
    }
}

template <class FT>
int bmt1m_impl(int mode, const typename P<FT>::params_1m* p, int64_t n, const FT* const (&in)[7], FT dt, int nsub, FT* const* out4,
               FT* const* src18, void* stream) {
    int st = validate_columns<FT, 7>(p, n, in);
    if (st) return st;
    if ((st = check_1m_options<FT>(p))) return st;
    if (out4 == nullptr) return cmh::fail(CUMICRO_E_NULL, "output pointer table is NULL");
    FT* o4[4] = {out4[0], out4[1], out4[2], out4[3]};
    // Verbose: any subset of the 4 + 18 columns (NULL = not wanted) — the stand-alone leaf methods of CM1 / MicrophysicsNonEq
    // (NEQ:110-224, CM1:352-1139) are single source-term columns of this kernel
    if ((st = require_outputs<FT, 4>(n, o4, mode == 1 ? 0 : 4))) return st;
    if (mode == 1 && src18 == nullptr) return cmh::fail(CUMICRO_E_NULL, "source-term pointer table is NULL");
    if (mode == 2) {
        if (!(dt > FT(0))) return cmh::fail(CUMICRO_E_ARG, "LinearizedAverage: dt must be > 0");
        if (nsub < 1) return cmh::fail(CUMICRO_E_ARG, "LinearizedAverage: nsub = %d must be >= 1", nsub);
    }
    cudaStream_t s = (cudaStream_t)stream;
    // the default exponent structure (quarter / eighth powers of λ⁻¹ only) runs the body that forms them by multiplication
    P<D>::params_1m wide;
    widen(*p, wide);
    bool std_exponents = make_1m_k<D>(wide, is_f32<FT>()).std_exponents != 0;
#ifdef CUMICRO_TUNING
    { static const char* g = getenv("CUMICRO_1M_GENERIC"); if (g && g[0] == '1') std_exponents = false; }
#endif
    return std_exponents ? bmt1m_launch<FT, true>(mode, p, n, in, dt, nsub, o4, src18, s)
                         : bmt1m_launch<FT, false>(mode, p, n, in, dt, nsub, o4, src18, s);
}

// ---- terminal velocities: (rho, q) -> v -------------------------------------------------------
// kind: 0 rain Blk1M, 1 snow Blk1M, 2 rain Chen2022, 3 snow Chen2022 (large ice),
//       4 cloud liquid Stokes (monodisperse), 5 cloud ice Chen2022 small ice (monodisperse)
struct TermVel1M : OneMBase {
    using FT = D;
    int kind;
    P<D>::vel_chen_rain chen_rain;
    P<D>::vel_chen_small_ice chen_small;
    P<D>::vel_chen_large_ice chen_large;
    P<D>::vel_stokes stokes;
    __device__ __forceinline__ void operator()(const D (&x)[2], D (&y)[1]) const {
        const FT e = tk.eps_n;
        const FT rho = x[0], q = x[1];
        const FT qp = clamp0_(q), rhop = clamp0_(rho);
        FT w = FT(0);
        if (kind == 0) {
            const FT v0 = sqrt_(FT(8.0 / 3) / p.vel_rain.C_drag * fmax_(p.vel_rain.rho_w / rho - FT(1), FT(0)) * p.vel_rain.grav * p.vel_rain.r0);
            w = terminal_velocity_blk1m<FT>(e, k.vt_rai_pref * v0, k.vt_rai_x, k.rai, log_full_(rhop * qp), k.log_n0_rai, q);
        } else if (kind == 1) {
            const FT L = log_full_(rhop * qp);
            const FT log_n0 = (q > e) ? fma_(p.snow.nu, log_full_(rho * fmax_(q, e)), k.log_mu_sno) : k.log_eps_numerics;
            w = terminal_velocity_blk1m<FT>(e, k.vt_sno_pref, k.vt_sno_x, k.sno, L, log_n0, q);
        } else if (kind == 2 || kind == 3) {
            FT lam, ll;
            if (kind == 2) {
                lambda_inverse(k.rai, log_full_(rhop * qp), k.log_n0_rai, lam, ll);
                FT aiu[3], bi[3], ciu[3];
                chen2022_vel_coeffs_rain<FT>(chen_rain, rho, aiu, bi, ciu);
                w = clamp0_(chen_exponential_pdf_sum3<FT, 3>(aiu, bi, ciu, FT(2) * lam));
            } else {
                const FT log_n0 = (q > e) ? fma_(p.snow.nu, log_full_(rho * fmax_(q, e)), k.log_mu_sno) : k.log_eps_numerics;
                lambda_inverse(k.sno, log_full_(rhop * qp), log_n0, lam, ll);
                // CO.Chen2022_vel_coeffs(::Chen2022VelTypeLargeIce, ρₐ, ρᵢ)          CO:326-349
                const FT ra = fmax_(rho, FT(0));
                const FT ri = p.snow.rho_i, l = log_full_(ri), sq = sqrt_(ri);
                const auto& c = chen_large;
                const FT Al = c.A[0] + c.A[1] * l + c.A[2] / (ri * sq);
                const FT Bl = exp_full_(c.B[0] + c.B[1] * (l * l) + c.B[2] * l);
                const FT Cl = exp_full_(c.C[0] + c.C[1] / l + c.C[2] / ri);
                const FT El = c.E[0] + c.E[1] * l * sq + c.E[2] * sq;
                const FT Fl = c.F[0] + c.F[1] * l - exp_full_(log_full_(-c.F[2]) - ri);
                const FT Gl = FT(1) / (c.G[0] + c.G[1] * l * sq + c.G[2] / sq);
                const FT Hl = c.H[0] + c.H[1] * (ri * ri) * sq + exp_full_(log_full_(-c.H[2]) - ri);
                const FT pa = pow_full_(ra, Al);
                const FT bi[2] = {Cl, Fl};
                const FT log1000 = FT(6.907755278982137);
                const FT aiu[2] = {Bl * pa * exp_full_(bi[0] * log1000), El * pa * exp_full_(Hl * ra) * exp_full_(bi[1] * log1000)};
                const FT ciu[2] = {FT(0), Gl * FT(1000)};
                const FT pk = pow_full_(p.snow.aspr_phi, p.snow.aspr_kappa);
                w = clamp0_(pk * chen_exponential_pdf_sum3<FT, 2>(aiu, bi, ciu, FT(2) * lam));
            }
            w = (q > e) ? w : FT(0);
        } else if (kind == 4) {
            const FT pref = FT(1.0 / 18) * (stokes.rho_w / rho - FT(1)) * stokes.grav / stokes.nu_air;
            const FT D = cbrt_full_(FT(6 / 3.141592653589793238462643383279502884L) * rho * qp / p.cloud_liquid.N_0 / p.cloud_liquid.rho_w);
            w = (q > e) ? pref * (D * D) : FT(0);
        } else {
            // CO.Chen2022_vel_coeffs(::Chen2022VelTypeSmallIce, ρₐ, ρᵢ)              CO:302-324
            const FT ra = fmax_(rho, FT(0));
            const FT ri = p.cloud_ice.rho_i, l = log_full_(ri), sq = sqrt_(ri);
            const auto& c = chen_small;
            const FT As = c.A[1] * (l * l) - c.A[2] * l + c.A[0];
            const FT Bs = FT(1) / (c.B[0] + c.B[1] * l + c.B[2] / sq);
            const FT Cs = c.C[0] + c.C[1] * exp_full_(c.C[2] * ri) + c.C[3] * sq;
            const FT Es = c.E[0] - c.E[1] * (l * l) + c.E[2] * sq;
            const FT Fs = -exp_full_(c.F[0] - c.F[1] * (l * l) + c.F[2] * l);
            const FT Gs = FT(1) / (c.G[0] + c.G[1] / l - c.G[2] * l / ri);
            const FT pa = pow_full_(ra, As);
            const FT b = Bs + ra * Cs;
            const FT u = exp_full_(b * FT(6.907755278982137));
            const FT D = cbrt_full_(FT(6 / 3.141592653589793238462643383279502884L) * rho * qp / p.cloud_ice.N_0 / p.cloud_ice.rho_i);
            const FT Db = pow_full_(D, b);
            const FT v = Es * pa * u * Db + Fs * pa * u * Db * exp_full_(-(Gs * FT(1000)) * D);   // Chen2022VelocityCurve
            w = (q > e) ? clamp0_(v) : FT(0);
        }
        y[0] = w;
    }
};

template <class FT>
int termvel_1m_impl(const typename P<FT>::params_1m* p, const void* vel, int kind, int64_t n, const FT* rho, const FT* q, FT* out,
                    void* stream) {
    const FT* in[2] = {rho, q};
    FT* o[1] = {out};
    int st = validate_columns<FT, 2>(p, n, in);
    if (st) return st;
    if ((st = require_outputs<FT, 1>(n, o, 1))) return st;
    if (kind < 0 || kind > 5) return cmh::fail(CUMICRO_E_OPTION, "termvel_1m: kind = %d (expected 0..5)", kind);
    if (kind >= 2 && vel == nullptr) return cmh::fail(CUMICRO_E_NULL, "termvel_1m: velocity parameter block is NULL");
    TermVel1M f = make_1m<FT, TermVel1M>(p);
    f.kind = kind;
    if (kind == 2) widen(*static_cast<const typename P<FT>::vel_chen_rain*>(vel), f.chen_rain);
    if (kind == 3) widen(*static_cast<const typename P<FT>::vel_chen_large_ice*>(vel), f.chen_large);
    if (kind == 4) widen(*static_cast<const typename P<FT>::vel_stokes*>(vel), f.stokes);
    if (kind == 5) widen(*static_cast<const typename P<FT>::vel_chen_small_ice*>(vel), f.chen_small);
    return launch_pointwise<FT, 2, 1, TermVel1M, 256, 2>(f, n, in, o, (cudaStream_t)stream, "termvel_1m launch");
}

}  // namespace

extern "C" {

#define CUMICRO_DEF_1M(SUF, FT)                                                                                        \
    int cumicro_bmt1m_inst_##SUF(const cumicro_params_1m_##SUF* p, int64_t n, const FT* rho, const FT* T,               \
                                 const FT* q_tot, const FT* q_lcl, const FT* q_icl, const FT* q_rai, const FT* q_sno,  \
                                 FT* const* out4, void* stream) {                                                      \
        const FT* in[7] = {rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno};                                                 \
        return bmt1m_impl<FT>(0, p, n, in, FT(0), 1, out4, nullptr, stream);                                           \
    }                                                                                                                  \
    int cumicro_bmt1m_verbose_##SUF(const cumicro_params_1m_##SUF* p, int64_t n, const FT* rho, const FT* T,            \
                                    const FT* q_tot, const FT* q_lcl, const FT* q_icl, const FT* q_rai,                \
                                    const FT* q_sno, FT* const* out4, FT* const* src18, void* stream) {                \
        const FT* in[7] = {rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno};                                                 \
        return bmt1m_impl<FT>(1, p, n, in, FT(0), 1, out4, src18, stream);                                             \
    }                                                                                                                  \
    int cumicro_bmt1m_linavg_##SUF(const cumicro_params_1m_##SUF* p, int64_t n, const FT* rho, const FT* T,             \
                                   const FT* q_tot, const FT* q_lcl, const FT* q_icl, const FT* q_rai,                 \
                                   const FT* q_sno, FT dt, int nsub, FT* const* out4, void* stream) {                  \
        const FT* in[7] = {rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno};                                                 \
        return bmt1m_impl<FT>(2, p, n, in, dt, nsub, out4, nullptr, stream);                                           \
    }

#define CUMICRO_DEF_TV1M(SUF, FT)                                                                                      \
    int cumicro_termvel_1m_##SUF(const cumicro_params_1m_##SUF* p, const void* vel, int kind, int64_t n, const FT* rho, \
                                 const FT* q, FT* out, void* stream) {                                                 \
        return termvel_1m_impl<FT>(p, vel, kind, n, rho, q, out, stream);                                              \
    }
CUMICRO_DEF_TV1M(f64, double)
CUMICRO_DEF_TV1M(f32, float)

CUMICRO_DEF_1M(f64, double)
CUMICRO_DEF_1M(f32, float)

}  // extern "C"

// ---- 0-moment scheme (BMT:658-680, src/Microphysics0M.jl:35-46): HBM-bound, 24-32 B/point ----------------------
namespace {
template <class FT, bool WITH_SAT> struct ZeroM {
    FT tau_precip, qc_0, S_0;
    __device__ __forceinline__ void operator()(const double (&x)[WITH_SAT ? 3 : 2], double (&y)[1]) const {
        // the Float32 method's arithmetic is Float32 (three operations: nothing to gain from widening)
        const FT q_lcl = cm::clamp0_((FT)x[0]), q_icl = cm::clamp0_((FT)x[1]);
        const FT thr = WITH_SAT ? S_0 * (FT)x[2] : qc_0;
        y[0] = (double)(-cm::clamp0_(q_lcl + q_icl - thr) / tau_precip);
    }
};
template <class FT, class PB>
int bmt0m_impl(const PB* p, int64_t n, const FT* q_lcl, const FT* q_icl, const FT* q_vap_sat, FT* out, void* stream) {
    FT* o[1] = {out};
    if (p == nullptr) return cmh::fail(CUMICRO_E_NULL, "parameter block is NULL");
    int st;
    if (q_vap_sat) {
        const FT* in[3] = {q_lcl, q_icl, q_vap_sat};
        if ((st = cm::validate_columns<FT, 3>(p, n, in))) return st;
        if ((st = cm::require_outputs<FT, 1>(n, o, 1))) return st;
        return cm::launch_pointwise<FT, 3, 1, ZeroM<FT, true>, 256, 4>(ZeroM<FT, true>{p->tau_precip, p->qc_0, p->S_0}, n, in, o,
                                                                      (cudaStream_t)stream, "bmt0m launch");
    }
    const FT* in[2] = {q_lcl, q_icl};
    if ((st = cm::validate_columns<FT, 2>(p, n, in))) return st;
    if ((st = cm::require_outputs<FT, 1>(n, o, 1))) return st;
    return cm::launch_pointwise<FT, 2, 1, ZeroM<FT, false>, 256, 4>(ZeroM<FT, false>{p->tau_precip, p->qc_0, p->S_0}, n, in, o,
                                                                   (cudaStream_t)stream, "bmt0m launch");
}
}  // namespace
extern "C" {
int cumicro_bmt0m_f64(const cumicro_params_0m_f64* p, int64_t n, const double* q_lcl, const double* q_icl, const double* q_vap_sat,
                      double* dq_tot_dt, void* stream) {
    return bmt0m_impl<double>(p, n, q_lcl, q_icl, q_vap_sat, dq_tot_dt, stream);
}
int cumicro_bmt0m_f32(const cumicro_params_0m_f32* p, int64_t n, const float* q_lcl, const float* q_icl, const float* q_vap_sat,
                      float* dq_tot_dt, void* stream) {
    return bmt0m_impl<float>(p, n, q_lcl, q_icl, q_vap_sat, dq_tot_dt, stream);
}
}  // extern "C"
