// kernels_diag.cu — cloud diagnostics (reference: src/CloudDiagnostics.jl:30-187): radar reflectivity and effective radius
// of the 1-moment and 2-moment (SB2006) size distributions, and the Liu & Hallett (1997) 1/3 power law.
// Leaf diagnostics next to the tendency path: same columns as the 2M kernel, one pass, both outputs.
#include <cuda_runtime.h>

#include <cmath>

#include "cm_launch.cuh"
#include "cm_sb2006.cuh"

namespace {
using namespace cm;
using D = double;

CM_DEV bool notvalid(D B) { return B == 0.0 || !isfinite(B); }

// (radar_reflectivity_2M, effective_radius_2M): x = (q_lcl, q_rai, N_lcl, N_rai, rho)        CloudDiagnostics.jl:60-116
struct Diag2M {
    P<D>::sb_pdf_c pdf_c;
    P<D>::sb_pdf_r pdf_r;
    D pi_rho_w, eps, eps_n;
    D C, C_23;                 // 4/3 π ρw and its 2/3 power
    D gc[3], gr[3], gc0, gr0;  // Γ((ν+1+n)/μ) for n = 2, 1, 2/3 and Γ((ν+1)/μ), cloud / rain
    D ec[3], er[3];            // -n/μ
    __device__ __forceinline__ void operator()(const D (&x)[5], D (&y)[2]) const {
        const D q_lcl = x[0], q_rai = x[1], N_lcl = x[2], N_rai = x[3], rho = x[4];
        // CM2.pdf_rain_parameters_mass: Br = cbrt(6 / xr_mean)                              CM2:141-146
        const RainPDF<D> rp = pdf_rain_parameters<D>(pdf_r, pi_rho_w, eps, q_rai, rho, N_rai);
        const D Br = cbrt_full_(6.0 / rp.xr_mean);   // xr_mean = 0 (no rain) -> Inf -> not valid
        // CM2.pdf_cloud_parameters_mass: Bc = exp(logB)                                      CM2:176-202
        const D safe_q = fmax_(q_lcl, eps), safe_N = fmax_(N_lcl, eps);
        const D logx = log_full_(rho * safe_q / safe_N);
        const D logB = -pdf_c.mu_c * (logx + pdf_c.loggamma_z1 - pdf_c.loggamma_z2);
        const bool cond = (N_lcl < eps) || (q_lcl < eps);
        const D Bc = cond ? INFINITY : exp_full_(logB);
        const bool nc = notvalid(Bc), nr = notvalid(Br);
        // DT.generalized_gamma_Mⁿ = N B^(-n/μ) Γ((ν+1+n)/μ) / Γ((ν+1)/μ)                      DT:109-112
        D Mc[3], Mr[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            Mc[i] = nc ? 0.0 : N_lcl * pow_full_(Bc, ec[i]) * gc[i] / gc0;
            Mr[i] = nr ? 0.0 : N_rai * pow_full_(Br, er[i]) * gr[i] / gr0;
        }
        const D Zc = nc ? 0.0 : Mc[0] / (C * C), Zr = nr ? 0.0 : Mr[0] / (C * C);
        y[0] = fmax_(-150.0, 10.0 * (log10(clamp0_(Zc + Zr)) - (-18.0)));
        const D M3 = (nc ? 0.0 : Mc[1] / C) + (nr ? 0.0 : Mr[1] / C);
        const D M2 = (nc ? 0.0 : Mc[2] / C_23) + (nr ? 0.0 : Mr[2] / C_23);
        y[1] = (M2 <= eps_n) ? 0.0 : M3 / M2;
    }
};

// radar_reflectivity_1M: x = (q_rai, rho)                                                   CloudDiagnostics.jl:30-42
struct Diag1M {
    D n0, inv_exp, r0_pow, denom, lam_floor;   // CM1.lambda_inverse pieces (CM1:126-152)
    D eps_n, c1em12, c1em3;
    __device__ __forceinline__ void operator()(const D (&x)[2], D (&y)[1]) const {
        const D q = clamp0_(x[0]), rho = clamp0_(x[1]);
        const D lam = fmax_(lam_floor, pow_full_(rho * q * r0_pow / denom, inv_exp));
        const D lam_mm = lam / c1em3;
        const D Z = 720.0 * (n0 * c1em12) * pow_full_(lam_mm, 7.0);
        y[0] = fmax_(-150.0, 10.0 * (log10(Z) - (-18.0) - 9.0));
    }
};

// effective_radius_Liu_Hallet_97: x = (rho, q_lcl, N_lcl, q_rai, N_rai)                      CloudDiagnostics.jl:132-165
template <bool FULL> struct ReffLH97 {
    D rho_w, eps_n, third, k_third;
    __device__ __forceinline__ void operator()(const D (&x)[FULL ? 5 : 2], D (&y)[1]) const {
        const D N = FULL ? x[2] + x[4] : 100.0 + 0.0;
        const D q = FULL ? x[1] + x[3] : x[1] + 0.0;
        const D r_vol = (N < eps_n) ? 0.0 : pow_full_((3.0 * q * x[0]) / (4.0 * 3.141592653589793 * rho_w * N), third);
        y[0] = r_vol / k_third;
    }
};

template <class FT> D eps_of() { return sizeof(FT) == 4 ? 1.1920928955078125e-07 : 2.220446049250313e-16; }
template <class FT> D epsn_of() { return sizeof(FT) == 4 ? 2.2737367544323206e-13 : 2.8126442852362996e-103; }

template <class FT, class PC, class PR>
int diag_2m_impl(const PC* pc, const PR* pr, int64_t n, const FT* q_lcl, const FT* q_rai, const FT* N_lcl, const FT* N_rai, const FT* rho,
                 FT* Z, FT* reff, void* stream) {
    if (pc == nullptr || pr == nullptr) return cmh::fail(CUMICRO_E_NULL, "size-distribution parameter block is NULL");
    const FT* in[5] = {q_lcl, q_rai, N_lcl, N_rai, rho};
    FT* out[2] = {Z, reff};
    int st;
    if ((st = validate_columns<FT, 5>(pc, n, in))) return st;
    if (n > 0 && Z == nullptr && reff == nullptr) return cmh::fail(CUMICRO_E_NULL, "both output columns are NULL");
    Diag2M f{};
    widen(*pc, f.pdf_c);
    widen(*pr, f.pdf_r);
    const D pi = 3.141592653589793238462643383279502884;
    f.pi_rho_w = pi * f.pdf_r.rho_w;
    f.eps = eps_of<FT>();
    f.eps_n = epsn_of<FT>();
    // 4 / 3 * π * ρw (CloudDiagnostics.jl:64, 99).  Float32 method: parameters and thresholds are the Float32 ones, literal
    // constants stay at Float64 precision (the result is judged against the method's exact-arithmetic value)
    f.C = 4.0 / 3 * 3.141592653589793 * f.pdf_r.rho_w;
    const D nm23 = 2.0 / 3;
    f.C_23 = std::pow(f.C, nm23);
    const D ns[3] = {2.0, 1.0, nm23};
    for (int i = 0; i < 3; ++i) {
        f.gc[i] = std::tgamma((f.pdf_c.nu_c + 1 + ns[i]) / f.pdf_c.mu_c);
        f.gr[i] = std::tgamma((f.pdf_r.nu_r + 1 + ns[i]) / f.pdf_r.mu_r);
        f.ec[i] = -ns[i] / f.pdf_c.mu_c;
        f.er[i] = -ns[i] / f.pdf_r.mu_r;
    }
    f.gc0 = std::tgamma((f.pdf_c.nu_c + 1) / f.pdf_c.mu_c);
    f.gr0 = std::tgamma((f.pdf_r.nu_r + 1) / f.pdf_r.mu_r);
    return launch_pointwise<FT, 5, 2, Diag2M, 256, 2>(f, n, in, out, (cudaStream_t)stream, "diag_2m launch");
}

template <class FT, class PB> int diag_1m_impl(const PB* p, int64_t n, const FT* q_rai, const FT* rho, FT* Z, void* stream) {
    const FT* in[2] = {q_rai, rho};
    FT* out[1] = {Z};
    int st;
    if ((st = validate_columns<FT, 2>(p, n, in))) return st;
    if ((st = require_outputs<FT, 1>(n, out, 1))) return st;
    const bool f32 = sizeof(FT) == 4;
    Diag1M f{};
    const auto& m = p->rain.mass;
    const D e = epsn_of<FT>();
    f.n0 = p->rain.n0;
    f.eps_n = e;
    f.inv_exp = 1.0 / ((D)m.me + (D)m.dm + 1.0);
    f.r0_pow = std::pow((D)m.r0, (D)m.me + (D)m.dm);
    f.denom = (D)m.chi_m * (D)m.m0 * std::max((D)p->rain.n0, e) * (D)m.gamma_coeff;
    f.lam_floor = (D)m.r0 * (f32 ? (D)1e-5f : 1e-5);   // a threshold: the method's own FT(1e-5) (CM1:151)
    f.c1em12 = 1e-12;
    f.c1em3 = 1e-3;
    return launch_pointwise<FT, 2, 1, Diag1M, 256, 2>(f, n, in, out, (cudaStream_t)stream, "diag_1m launch");
}

template <class FT>
int diag_lh97_impl(FT rho_w, int64_t n, const FT* rho, const FT* q_lcl, const FT* N_lcl, const FT* q_rai, const FT* N_rai, FT* reff, void* stream) {
    FT* out[1] = {reff};
    const D third = 1.0 / 3, k = 0.8;
    int st;
    const int dummy = 0;
    if ((st = require_outputs<FT, 1>(n, out, 1))) return st;
    const bool full = N_lcl || q_rai || N_rai;
    if (full) {
        const FT* in[5] = {rho, q_lcl, N_lcl, q_rai, N_rai};
        if ((st = validate_columns<FT, 5>(&dummy, n, in))) return st;
        ReffLH97<true> f{(D)rho_w, epsn_of<FT>(), third, std::pow(k, third)};
        return launch_pointwise<FT, 5, 1, ReffLH97<true>, 256, 2>(f, n, in, out, (cudaStream_t)stream, "diag_reff_lh97 launch");
    }
    const FT* in[2] = {rho, q_lcl};
    if ((st = validate_columns<FT, 2>(&dummy, n, in))) return st;
    ReffLH97<false> f{(D)rho_w, epsn_of<FT>(), third, std::pow(k, third)};
    return launch_pointwise<FT, 2, 1, ReffLH97<false>, 256, 2>(f, n, in, out, (cudaStream_t)stream, "diag_reff_lh97 launch");
}
}  // namespace

extern "C" {
#define CUMICRO_DEF_DIAG(SUF, FT)                                                                                                          \
    int cumicro_diag_2m_##SUF(const cumicro_sb_pdf_c_##SUF* pdf_c, const cumicro_sb_pdf_r_##SUF* pdf_r, int64_t n, const FT* q_lcl,         \
                              const FT* q_rai, const FT* N_lcl, const FT* N_rai, const FT* rho, FT* Z, FT* r_eff, void* stream) {          \
        return diag_2m_impl<FT>(pdf_c, pdf_r, n, q_lcl, q_rai, N_lcl, N_rai, rho, Z, r_eff, stream);                                       \
    }                                                                                                                                      \
    int cumicro_diag_1m_##SUF(const cumicro_params_1m_##SUF* p, int64_t n, const FT* q_rai, const FT* rho, FT* Z, void* stream) {          \
        return diag_1m_impl<FT>(p, n, q_rai, rho, Z, stream);                                                                              \
    }                                                                                                                                      \
    int cumicro_diag_reff_lh97_##SUF(FT rho_w, int64_t n, const FT* rho, const FT* q_lcl, const FT* N_lcl, const FT* q_rai,                \
                                     const FT* N_rai, FT* r_eff, void* stream) {                                                           \
        return diag_lh97_impl<FT>(rho_w, n, rho, q_lcl, N_lcl, q_rai, N_rai, r_eff, stream);                                               \
    }
CUMICRO_DEF_DIAG(f64, double)
CUMICRO_DEF_DIAG(f32, float)
}  // extern "C"
