// kernels_2m.cu — fused 2-moment (Seifert-Beheng 2006) warm-rain tendency kernels,
// the 2-moment terminal velocities, and their C-ABI entry points (include/cumicro.h).
//
// Layout: structure-of-arrays columns of length n in HBM; data movement is
// cm_launch.cuh's streaming kernel (one 128-bit load per input column and one
// 128-bit streaming store per output column per thread item).  The parameter block
// and the host-derived constants travel in kernel-parameter (constant-bank) space.
#include <cmath>
#include <cstdlib>
#include <limits>

#include "cm_hostpipe.cuh"
#include "cm_launch.cuh"
#include "cm_sb2006.cuh"

namespace {

using namespace cm;

// ---- BMT:820-854: 7 columns in, 4 tendencies out --------------------------------------
// (NIN = 8 adds the optional q_ice column that the reference's warm-only method
// accepts and forwards to the thermodynamics, BMT:823,836,843.)
template <class FT, int NIN = 7> struct Warm2MFused {
    typename P<FT>::params_2m_warm p;
    ThermoK<FT> tk;
    SB2006K<FT> sk;
    __device__ __forceinline__ void operator()(const FT (&x)[NIN], FT (&y)[4]) const {
        const FT q_ice = (NIN == 8) ? fmax_(FT(0), x[NIN - 1]) : FT(0);
        Warm2M<FT> o = warm_rain_tendencies_2m<FT>(p, tk, sk, x[0], x[1], x[2], x[3], x[4], x[5], x[6], q_ice);
        y[0] = o.dq_lcl_dt;
        y[1] = o.dn_lcl_dt;
        y[2] = o.dq_rai_dt;
        y[3] = o.dn_rai_dt;
    }
};

// ---- the 15 SB2006 process rates one by one (leaf API) ----------------------------------
template <class FT> struct Warm2MLeaves {
    typename P<FT>::params_2m_warm p;
    ThermoK<FT> tk;
    SB2006K<FT> sk;
    __device__ __forceinline__ void operator()(const FT (&x)[7], FT (&y)[CUMICRO_SB2006_NLEAF]) const {
        Warm2M<FT> o = warm_rain_tendencies_2m<FT>(p, tk, sk, x[0], x[1], x[2], x[3], x[4], x[5], x[6], FT(0));
#pragma unroll
        for (int k = 0; k < CUMICRO_SB2006_NLEAF; ++k) y[k] = o.leaf[k];
    }
};

template <class FT> int check_2m_options(const typename P<FT>::params_2m_warm* p) {
    if (p->sb.pdf_r.limited != 0 && p->sb.pdf_r.limited != 1)
        return cmh::fail(CUMICRO_E_OPTION, "sb.pdf_r.limited = %d (expected 0 or 1)", (int)p->sb.pdf_r.limited);
    return CUMICRO_OK;
}

template <class FT>
int bmt2m_warm_impl(const typename P<FT>::params_2m_warm* p, int64_t n, const FT* rho, const FT* T, const FT* q_tot,
                    const FT* q_lcl, const FT* n_lcl, const FT* q_rai, const FT* n_rai, const FT* q_ice,
                    FT* dq_lcl_dt, FT* dn_lcl_dt, FT* dq_rai_dt, FT* dn_rai_dt, FT* const* zero4, void* stream) {
    const FT* in[7] = {rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai};
    FT* out[4] = {dq_lcl_dt, dn_lcl_dt, dq_rai_dt, dn_rai_dt};
    int st = validate_columns<FT, 7>(p, n, in);
    if (st) return st;
    if ((st = check_2m_options<FT>(p))) return st;
    if ((st = require_outputs<FT, 4>(n, out, 4))) return st;
    cudaStream_t s = (cudaStream_t)stream;
    if (zero4 && n > 0)
        for (int k = 0; k < 4; ++k)
            if (zero4[k]) {
                st = cmh::cuda_status(cudaMemsetAsync(zero4[k], 0, sizeof(FT) * (size_t)n, s), "cudaMemsetAsync");
                if (st) return st;
            }
    if (q_ice != nullptr) {
        const FT* in8[8] = {rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice};
        Warm2MFused<FT, 8> f8{*p, make_thermo_k<FT>(p->tps), make_sb2006_k<FT>(p->sb, p->aps)};
        return launch_pointwise<FT, 8, 4, Warm2MFused<FT, 8>, 128, 8, false>(f8, n, in8, out, s, "bmt2m_warm kernel launch");
    }
    Warm2MFused<FT> f{*p, make_thermo_k<FT>(p->tps), make_sb2006_k<FT>(p->sb, p->aps)};
#ifdef CUMICRO_TUNING
    {   // launch-shape exploration (tools/tune_2m.py); not compiled into the product build
        static const char* ev = getenv("CUMICRO_2M_VARIANT");
        const int v = ev ? atoi(ev) : 0;
        const char* w = "bmt2m_warm kernel launch";
        switch (v) {
            case 1: return launch_pointwise<FT, 7, 4, Warm2MFused<FT>, 256, 2>(f, n, in, out, s, w);
            case 2: return launch_pointwise<FT, 7, 4, Warm2MFused<FT>, 128, 3>(f, n, in, out, s, w);
            case 3: return launch_pointwise<FT, 7, 4, Warm2MFused<FT>, 128, 4>(f, n, in, out, s, w);
            case 4: return launch_pointwise<FT, 7, 4, Warm2MFused<FT>, 128, 5>(f, n, in, out, s, w);
            case 5: return launch_pointwise<FT, 7, 4, Warm2MFused<FT>, 128, 6>(f, n, in, out, s, w);
            case 6: return launch_pointwise<FT, 7, 4, Warm2MFused<FT>, 128, 8>(f, n, in, out, s, w);
            case 7: return launch_pointwise<FT, 7, 4, Warm2MFused<FT>, 64, 8>(f, n, in, out, s, w);
            case 8: return launch_pointwise<FT, 7, 4, Warm2MFused<FT>, 64, 12>(f, n, in, out, s, w);
            default: break;
        }
    }
#endif
    return launch_pointwise<FT, 7, 4, Warm2MFused<FT>, 128, 8, false>(f, n, in, out, s, "bmt2m_warm kernel launch");
}

template <class FT>
int sb2006_leaves_impl(const typename P<FT>::params_2m_warm* p, int64_t n, const FT* rho, const FT* T,
                       const FT* q_tot, const FT* q_lcl, const FT* n_lcl, const FT* q_rai, const FT* n_rai,
                       FT* const* out_tbl, void* stream) {
    const FT* in[7] = {rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai};
    int st = validate_columns<FT, 7>(p, n, in);
    if (st) return st;
    if ((st = check_2m_options<FT>(p))) return st;
    if (out_tbl == nullptr) return cmh::fail(CUMICRO_E_NULL, "leaf pointer table is NULL");
    FT* out[CUMICRO_SB2006_NLEAF];
    for (int k = 0; k < CUMICRO_SB2006_NLEAF; ++k) out[k] = out_tbl[k];
    Warm2MLeaves<FT> f{*p, make_thermo_k<FT>(p->tps), make_sb2006_k<FT>(p->sb, p->aps)};
    return launch_pointwise<FT, 7, CUMICRO_SB2006_NLEAF, Warm2MLeaves<FT>, 256, 1>(f, n, in, out, (cudaStream_t)stream,
                                                                                  "sb2006_leaves kernel launch");
}

template <class FT>
int bmt2m_warm_host_impl(const typename P<FT>::params_2m_warm* p, int64_t n, const FT* rho, const FT* T,
                         const FT* q_tot, const FT* q_lcl, const FT* n_lcl, const FT* q_rai, const FT* n_rai,
                         FT* dq_lcl_dt, FT* dn_lcl_dt, FT* dq_rai_dt, FT* dn_rai_dt, int64_t chunk) {
    const FT* in[7] = {rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai};
    FT* out[4] = {dq_lcl_dt, dn_lcl_dt, dq_rai_dt, dn_rai_dt};
    int st = validate_columns<FT, 7>(p, n, in);
    if (st) return st;
    if ((st = check_2m_options<FT>(p))) return st;
    if ((st = require_outputs<FT, 4>(n, out, 4))) return st;
    Warm2MFused<FT> f{*p, make_thermo_k<FT>(p->tps), make_sb2006_k<FT>(p->sb, p->aps)};
    return host_pipeline<FT, 7, 4>(n, in, out, chunk,
                                   [&](int64_t m, const FT* const(&din)[7], FT* const(&dout)[4], cudaStream_t s) {
                                       return launch_pointwise<FT, 7, 4, Warm2MFused<FT>, 128, 8, false>(
                                           f, m, din, dout, s, "bmt2m_warm (host pipeline) kernel launch");
                                   });
}

// ---- terminal velocities: (q, rho, N) -> (vt0, vt1) ---------------------------------------
template <class FT> struct RainVelSB {
    typename P<FT>::sb_pdf_r pdf_r;
    typename P<FT>::vel_sb2006 vel;
    FT pi_rho_w;
    __device__ __forceinline__ void operator()(const FT (&x)[3], FT (&y)[2]) const {
        rain_terminal_velocity_sb<FT>(pdf_r, vel, pi_rho_w, x[0], x[1], x[2], y[0], y[1]);
    }
};
template <class FT> struct RainVelChen {
    typename P<FT>::sb_pdf_r pdf_r;
    typename P<FT>::vel_chen_rain vel;
    FT pi_rho_w;
    __device__ __forceinline__ void operator()(const FT (&x)[3], FT (&y)[2]) const {
        rain_terminal_velocity_chen<FT>(pdf_r, vel, pi_rho_w, x[0], x[1], x[2], y[0], y[1]);
    }
};
template <class FT> struct CloudVel {
    typename P<FT>::sb_pdf_c pdf_c;
    typename P<FT>::vel_stokes vel;
    FT pref0;
    FT gratio[2];
    __device__ __forceinline__ void operator()(const FT (&x)[3], FT (&y)[2]) const {
        cloud_terminal_velocity<FT>(pdf_c, vel, pref0, gratio, x[0], x[1], x[2], y[0], y[1]);
    }
};

template <class FT, class F>
int termvel_impl(const void* p1, const void* p2, const F& f, int64_t n, const FT* q, const FT* rho, const FT* N, FT* vt0,
                 FT* vt1, void* stream, const char* what) {
    const FT* in[3] = {q, rho, N};
    FT* out[2] = {vt0, vt1};
    if (p2 == nullptr) return cmh::fail(CUMICRO_E_NULL, "velocity parameter block is NULL");
    int st = validate_columns<FT, 3>(p1, n, in);
    if (st) return st;
    if ((st = require_outputs<FT, 2>(n, out, 2))) return st;
    return launch_pointwise<FT, 3, 2, F, 256, 2>(f, n, in, out, (cudaStream_t)stream, what);
}

template <class FT> FT pi_rho_w_of(const typename P<FT>::sb_pdf_r* pdf) {
    return pdf ? FT(3.141592653589793238462643383279502884L) * pdf->rho_w : FT(0);
}

template <class FT>
CloudVel<FT> make_cloud_vel(const typename P<FT>::sb_pdf_c* pdf, const typename P<FT>::vel_stokes* vel) {
    CloudVel<FT> f{};
    if (!pdf || !vel) return f;
    f.pdf_c = *pdf;
    f.vel = *vel;
    const FT pi = FT(3.141592653589793238462643383279502884L);
    const FT t = FT(6) / vel->rho_w / pi;
    f.pref0 = FT(1.0 / 18) * std::cbrt(t * t) * vel->grav / vel->nu_air;
    const FT z = (pdf->nu_c + 1) / pdf->mu_c;
    f.gratio[0] = std::tgamma((pdf->nu_c + 1 + FT(2.0 / 3)) / pdf->mu_c) / std::tgamma(z);
    f.gratio[1] = std::tgamma((pdf->nu_c + 1 + FT(5.0 / 3)) / pdf->mu_c) / std::tgamma(z);
    return f;
}

}  // namespace

extern "C" {

#define CUMICRO_DEF_2M(SUF, FT)                                                                                        \
    int cumicro_bmt2m_warm_##SUF(const cumicro_params_2m_warm_##SUF* p, int64_t n, const FT* rho, const FT* T,          \
                                 const FT* q_tot, const FT* q_lcl, const FT* n_lcl, const FT* q_rai, const FT* n_rai,  \
                                 const FT* q_ice, FT* dq_lcl_dt, FT* dn_lcl_dt, FT* dq_rai_dt, FT* dn_rai_dt,          \
                                 FT* const* zero4, void* stream) {                                                     \
        return bmt2m_warm_impl<FT>(p, n, rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, dq_lcl_dt, dn_lcl_dt,       \
                                   dq_rai_dt, dn_rai_dt, zero4, stream);                                               \
    }                                                                                                                  \
    int cumicro_bmt2m_warm_host_##SUF(const cumicro_params_2m_warm_##SUF* p, int64_t n, const FT* rho, const FT* T,     \
                                      const FT* q_tot, const FT* q_lcl, const FT* n_lcl, const FT* q_rai,              \
                                      const FT* n_rai, FT* dq_lcl_dt, FT* dn_lcl_dt, FT* dq_rai_dt, FT* dn_rai_dt,     \
                                      int64_t chunk) {                                                                 \
        return bmt2m_warm_host_impl<FT>(p, n, rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, dq_lcl_dt, dn_lcl_dt,         \
                                        dq_rai_dt, dn_rai_dt, chunk);                                                  \
    }                                                                                                                  \
    int cumicro_sb2006_leaves_##SUF(const cumicro_params_2m_warm_##SUF* p, int64_t n, const FT* rho, const FT* T,       \
                                    const FT* q_tot, const FT* q_lcl, const FT* n_lcl, const FT* q_rai,                \
                                    const FT* n_rai, FT* const* out, void* stream) {                                   \
        return sb2006_leaves_impl<FT>(p, n, rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, out, stream);                   \
    }                                                                                                                  \
    int cumicro_termvel_2m_rain_sb_##SUF(const cumicro_sb_pdf_r_##SUF* pdf_r, const cumicro_vel_sb2006_##SUF* vel,      \
                                         int64_t n, const FT* q_rai, const FT* rho, const FT* N_rai, FT* vt0, FT* vt1, \
                                         void* stream) {                                                               \
        RainVelSB<FT> f{};                                                                                             \
        if (pdf_r && vel) f = RainVelSB<FT>{*pdf_r, *vel, pi_rho_w_of<FT>(pdf_r)};                                     \
        return termvel_impl<FT>(pdf_r, vel, f, n, q_rai, rho, N_rai, vt0, vt1, stream, "termvel_2m_rain_sb launch");   \
    }                                                                                                                  \
    int cumicro_termvel_2m_rain_chen_##SUF(const cumicro_sb_pdf_r_##SUF* pdf_r,                                         \
                                           const cumicro_vel_chen_rain_##SUF* vel, int64_t n, const FT* q_rai,         \
                                           const FT* rho, const FT* N_rai, FT* vt0, FT* vt1, void* stream) {           \
        RainVelChen<FT> f{};                                                                                           \
        if (pdf_r && vel) f = RainVelChen<FT>{*pdf_r, *vel, pi_rho_w_of<FT>(pdf_r)};                                   \
        return termvel_impl<FT>(pdf_r, vel, f, n, q_rai, rho, N_rai, vt0, vt1, stream,                                 \
                                "termvel_2m_rain_chen launch");                                                        \
    }                                                                                                                  \
    int cumicro_termvel_2m_cloud_##SUF(const cumicro_sb_pdf_c_##SUF* pdf_c, const cumicro_vel_stokes_##SUF* vel,        \
                                       int64_t n, const FT* q_lcl, const FT* rho, const FT* N_lcl, FT* vt0, FT* vt1,   \
                                       void* stream) {                                                                 \
        CloudVel<FT> f = make_cloud_vel<FT>(pdf_c, vel);                                                               \
        return termvel_impl<FT>(pdf_c, vel, f, n, q_lcl, rho, N_lcl, vt0, vt1, stream, "termvel_2m_cloud launch");     \
    }

CUMICRO_DEF_2M(f64, double)
CUMICRO_DEF_2M(f32, float)

}  // extern "C"
