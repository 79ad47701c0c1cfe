// kernels_2m_alt.cu — the alternative 2-moment closures of Wood (2005): Khairoutdinov & Kogan 2000,
// Beheng 1994, Tripoli & Cotton 1980, Liu & Daum 2004 autoconversion and accretion
// (reference: src/Microphysics2M.jl:920-1002; goldens test/gpu_tests.jl:795-818).
// Leaf rates, not on the fused BMT path: three columns in, one out (32 B/point in Float64), full-range libm
// powers because q_lcl = 0 and the smooth-transition limits are legal inputs.
#include <cuda_runtime.h>

#include <cmath>

#include "cm_1m.cuh"
#include "cm_launch.cuh"

namespace {
using namespace cm;
using D = double;

// CO.logistic_function (src/Common.jl:124-138)
CM_DEV D logistic_function(D e, D x, D x_0, D k) {
    x = clamp0_(x);
    const D x_safe = fmax_(x, e), x0_safe = fmax_(x_0, e);
    const D z = k * (x_safe / x0_safe - x0_safe / x_safe);
    const D result = exp_full_(-log1pexp_(-z));
    return (x < e) ? 0.0 : ((x_0 < e) ? 1.0 : result);
}
CM_DEV D heaviside(D x) { return (x > 0.0) ? 1.0 : 0.0; }  // CO.heaviside (src/Common.jl:107-109)

struct Alt2M {
    cumicro_params_2m_alt_f64 p;
    D eps, eps_n;   // UT.ϵ_numerics_2M_M, UT.ϵ_numerics of the method's float type
    int what, smooth;
    // acnv (what <= 3): x = (q_lcl, rho, N_d); accretion (what >= 4): x = (q_lcl, q_rai, rho)
    __device__ __forceinline__ void operator()(const D (&x)[3], D (&y)[1]) const {
        D q_lcl = x[0];
        D r = 0.0;
        switch (what) {
            case 0: {  // CM2:920-924
                q_lcl = clamp0_(q_lcl);
                r = p.kk_acnv_A * pow_full_(q_lcl, p.kk_acnv_a) * pow_full_(x[2], p.kk_acnv_b) * pow_full_(x[1], p.kk_acnv_c);
                break;
            }
            case 1: {  // CM2:925-937
                q_lcl = clamp0_(q_lcl);
                const D rho = x[1], N_d = x[2];
                D d;
                if (smooth) {
                    const D lo = logistic_function(eps_n, N_d, p.b_acnv_N_0, p.b_acnv_k);
                    const D hi = 1.0 - lo;
                    d = lo * p.b_acnv_d_low + hi * p.b_acnv_d_high;
                } else {
                    d = (N_d >= p.b_acnv_N_0) ? p.b_acnv_d_low : p.b_acnv_d_high;
                }
                r = p.b_acnv_C * pow_full_(d, p.b_acnv_a) * pow_full_(q_lcl * rho, p.b_acnv_b) * pow_full_(N_d, p.b_acnv_c) / rho;
                break;
            }
            case 2: {  // CM2:938-947
                q_lcl = clamp0_(q_lcl);
                const D rho = x[1], N_d = x[2];
                const D thr = p.tc_acnv_m0_liq_coeff * N_d / rho * pow_full_(p.tc_acnv_r_0, p.tc_acnv_me_liq);
                const D o = smooth ? logistic_function(eps_n, q_lcl, thr, p.tc_acnv_k) : heaviside(q_lcl - thr);
                r = p.tc_acnv_D * pow_full_(q_lcl, p.tc_acnv_a) * pow_full_(N_d, p.tc_acnv_b) * o;
                break;
            }
            case 3: {  // CM2:948-969
                const D rho = x[1], N_d = x[2];
                if (q_lcl <= eps) break;
                const D r_vol = cbrt_full_(3.0 * q_lcl * rho / 4.0 / 3.141592653589793 / p.ld_rho_w / N_d) * 1000000.0;
                const D b6 = cbrt_full_((r_vol + 3.0) / r_vol);
                const D b2 = b6 * b6;
                const D E = p.ld_E_0 * (b2 * b2 * b2);
                const D R6 = b6 * r_vol;
                const D R6C = p.ld_R_6C_0 / cbrt_full_(sqrt_(q_lcl * rho)) / sqrt_(R6);
                const D o = smooth ? logistic_function(eps_n, R6, R6C, p.ld_k) : heaviside(R6 - R6C);
                const D L = q_lcl * rho;
                r = E * (L * L * L) / N_d / rho * o;
                break;
            }
            case 4: {  // CM2:985-990
                q_lcl = clamp0_(q_lcl);
                const D q_rai = clamp0_(x[1]);
                r = p.kk_accr_A * pow_full_(q_lcl * q_rai, p.kk_accr_a) * pow_full_(x[2], p.kk_accr_b);
                break;
            }
            case 5: {  // CM2:992-997
                q_lcl = clamp0_(q_lcl);
                const D q_rai = clamp0_(x[1]);
                r = p.b_accr_A * q_lcl * x[2] * q_rai;
                break;
            }
            default: {  // 6: CM2:999-1005
                q_lcl = clamp0_(q_lcl);
                const D q_rai = clamp0_(x[1]);
                r = p.tc_accr_A * q_lcl * q_rai;
                break;
            }
        }
        y[0] = r;
    }
};

template <class FT, class PB>
int alt2m_impl(const PB* p, int what, int smooth, int64_t n, const FT* q_lcl, const FT* q_rai, const FT* rho, const FT* N_d, FT* out,
               void* stream) {
    if (p == nullptr) return cmh::fail(CUMICRO_E_NULL, "parameter block is NULL");
    if (what < 0 || what > 6) return cmh::fail(CUMICRO_E_ARG, "2m_alt: unknown closure %d (0..6)", what);
    if (what == 6 && rho == nullptr) rho = q_lcl;  // accretion(::TC1980, q_lcl, q_rai) takes no density
    const FT* in[3];
    if (what <= 3) { in[0] = q_lcl; in[1] = rho; in[2] = N_d; }
    else { in[0] = q_lcl; in[1] = q_rai; in[2] = rho; }
    FT* o[1] = {out};
    int st;
    if ((st = validate_columns<FT, 3>(p, n, in))) return st;
    if ((st = require_outputs<FT, 1>(n, o, 1))) return st;
    Alt2M f{};
    widen(*p, f.p);
    const bool f32 = sizeof(FT) == 4;
    f.eps = f32 ? 1.1920928955078125e-07 : 2.220446049250313e-16;
    f.eps_n = f32 ? 2.2737367544323206e-13 : 2.8126442852362996e-103;
    f.what = what;
    f.smooth = smooth != 0;
    return launch_pointwise<FT, 3, 1, Alt2M, 256, 2>(f, n, in, o, (cudaStream_t)stream, "2m_alt launch");
}
}  // namespace

extern "C" {
int cumicro_2m_alt_f64(const cumicro_params_2m_alt_f64* p, int what, int smooth_transition, int64_t n, const double* q_lcl,
                       const double* q_rai, const double* rho, const double* N_d, double* out, void* stream) {
    return alt2m_impl<double>(p, what, smooth_transition, n, q_lcl, q_rai, rho, N_d, out, stream);
}
int cumicro_2m_alt_f32(const cumicro_params_2m_alt_f32* p, int what, int smooth_transition, int64_t n, const float* q_lcl,
                       const float* q_rai, const float* rho, const float* N_d, float* out, void* stream) {
    return alt2m_impl<float>(p, what, smooth_transition, n, q_lcl, q_rai, rho, N_d, out, stream);
}
}  // extern "C"
