// kernels_fused.cu — the "ClimaAtmos-scale" fused kernel of BASELINE config 5:
// 1-moment tendencies + 2-moment warm-rain tendencies + ice-nucleation rates (+ ARG2000
// activated number) from ONE read of the 11 state columns, all 11 outputs written once,
// and the domain diagnostics (precipitation production, activated number) reduced in the
// kernel epilogue: warp shuffles -> one partial per block -> a fixed-order second pass, so the
// sums are bit-reproducible and no output column is re-read.  The cross-GPU sum of the
// per-GPU diagnostics is the only collective of the whole path (NCCL all-reduce of
// CUMICRO_NDIAG doubles, done by the host binding).
#include <cmath>
#include <limits>

#include "cm_1m.cuh"
#include "cm_hostpipe.cuh"
#include "cm_icenuc.cuh"
#include "cm_launch.cuh"
#include "cm_sb2006.cuh"

namespace {

using namespace cm;
using D = double;
template <class FT> constexpr bool is_f32() { return sizeof(FT) == 4; }

constexpr int NIN = 11;   // rho, T, p, w, q_tot, q_lcl, q_icl, q_rai, q_sno, n_lcl, n_rai
constexpr int NOUT = 11;  // 1M: dq_lcl, dq_icl, dq_rai, dq_sno | 2M: dq_lcl, dn_lcl, dq_rai, dn_rai | J_dep, J_ABIFM, J_hom
constexpr int NDIAG = CUMICRO_NDIAG;
constexpr int BLOCK = 128;
#ifndef CUMICRO_FUSED_MINB
#define CUMICRO_FUSED_MINB 4
#endif

struct FusedParams {
    P<D>::params_1m p1;
    P<D>::params_2m_warm p2;
    cumicro_params_icenuc_f64 p3;
    ThermoK<D> tk;
    OneMK<D> k1;
    SB2006K<D> k2;
    ArgK<D> k3;
    int with_activation;
};

template <class FT> struct FusedArgs {
    FusedParams f;
    const FT* in[NIN];
    FT* out[NOUT];
    double* partials;   // [gridDim.x][NDIAG]
    int64_t n;
};

template <class FT>
__global__ void __launch_bounds__(BLOCK, CUMICRO_FUSED_MINB) fused_kernel(const __grid_constant__ FusedArgs<FT> a) {
    math_tables_init<BLOCK>();
    const FusedParams& f = a.f;
    double diag[NDIAG];
#pragma unroll
    for (int k = 0; k < NDIAG; ++k) diag[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * BLOCK;
    for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < a.n; i += stride) {
        double x[NIN];
#pragma unroll
        for (int c = 0; c < NIN; ++c) x[c] = (double)__ldg(a.in[c] + i);
        const double rho = x[0], T = x[1], pr = x[2], w = x[3], q_tot = x[4], q_lcl = x[5], q_icl = x[6], q_rai = x[7],
                     q_sno = x[8], n_lcl = x[9], n_rai = x[10];
        double y[NOUT];
        // 1-moment tendencies                                         BMT:505-514
        {
            const Src1M<D> r = microphysics_source_terms_1m<D>(f.p1, f.tk, f.k1, rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno);
            double t[4];
            aggregate_tendencies_1m<D>(r, t);
#pragma unroll
            for (int k = 0; k < 4; ++k) y[k] = t[k];
        }
        // 2-moment warm rain (cloud ice seen by the thermodynamics = q_icl + q_sno)   BMT:820-854
        {
            const Warm2M<D> o = warm_rain_tendencies_2m<D>(f.p2, f.tk, f.k2, rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai,
                                                           fmax_(0.0, q_icl) + fmax_(0.0, q_sno));
            y[4] = o.dq_lcl_dt;
            y[5] = o.dn_lcl_dt;
            y[6] = o.dq_rai_dt;
            y[7] = o.dn_rai_dt;
        }
        // ice-nucleation rates (+ ARG2000 activated number)              IN:92-134, 557-584; AA:138-273
        double n_act = 0.0;
        {
            double da_w;
            if (f.with_activation) {
                const ArgOut o = arg2000<false>(f.p3, f.tk, f.k3, T, pr, w, q_tot, q_lcl + q_rai, q_icl + q_sno, rho * n_lcl, 0.0);
                da_w = o.da_w;
#pragma unroll
                for (int m = 0; m < kMaxModes; ++m)
                    if (m < f.p3.n_modes) n_act += o.N_act[m];
            } else {
                const TempState<D> ts = temp_state(f.tk, T);
                const D pl = p_sat_liq(f.tk, ts);
                const D Rm = f.p3.tps.R_d * (1.0 + (f.k3.Rv_over_Rd - 1.0) * q_tot - f.k3.Rv_over_Rd * (q_lcl + q_rai + q_icl + q_sno));
                const D p_v = (q_tot - (q_lcl + q_rai) - (q_icl + q_sno)) * (pr / (Rm * T)) * f.tk.R_v * T;
                da_w = p_v / pl - p_sat_ice(f.tk, ts) / pl;
            }
            bool err = false;
            y[8] = deposition_J<D>(f.p3.dust, da_w, f.k3.ln10);
            y[9] = ABIFM_J<D>(f.p3.dust, da_w, f.k3.ln10);
            y[10] = f.p3.hom_linear ? homogeneous_J_linear<D>(f.p3.koop, da_w, f.k3.ln10)
                                    : homogeneous_J_cubic<D>(f.p3.koop, da_w, f.k3.ln10, err);
            if (err) y[10] = __longlong_as_double(0x7ff8000000000000LL);
        }
#pragma unroll
        for (int c = 0; c < NOUT; ++c)
            if (a.out[c]) __stcs(a.out[c] + i, (FT)y[c]);
        // diagnostics of this point (the rounded stored values are NOT used: Float64 sums)
        diag[0] += rho * (y[2] + y[3]);   // 1M precipitation production  Σ ρ (dq_rai + dq_sno)   [kg m^-3 s^-1]
        diag[1] += rho * y[6];            // 2M rain production           Σ ρ dq_rai
        diag[2] += n_act;                 // activated aerosol number     Σ N_act                  [m^-3]
        diag[3] += 1.0;                   // points
    }
    // block reduction: shuffle within warps, then across the (BLOCK/32) warps through shared memory
    __shared__ double red[BLOCK / 32][NDIAG];
#pragma unroll
    for (int k = 0; k < NDIAG; ++k) {
        double v = diag[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < NDIAG) {
        double v = 0.0;
#pragma unroll
        for (int wv = 0; wv < BLOCK / 32; ++wv) v += red[wv][threadIdx.x];
        a.partials[(size_t)blockIdx.x * NDIAG + threadIdx.x] = v;
    }
}

// fixed-order second pass: one block, thread k sums partial k of every block in block order
__global__ void fused_diag_finish(const double* partials, int n_blocks, double* diag) {
    const int k = threadIdx.x;
    if (k >= NDIAG) return;
    double v = 0.0;
    for (int b = 0; b < n_blocks; ++b) v += partials[(size_t)b * NDIAG + k];
    diag[k] = v;
}

template <class FT> struct PF;
template <> struct PF<double> { using p1 = cumicro_params_1m_f64; using p2 = cumicro_params_2m_warm_f64; using p3 = cumicro_params_icenuc_f64; };
template <> struct PF<float> { using p1 = cumicro_params_1m_f32; using p2 = cumicro_params_2m_warm_f32; using p3 = cumicro_params_icenuc_f32; };

template <class FT>
int fused_impl(const typename PF<FT>::p1* p1, const typename PF<FT>::p2* p2, const typename PF<FT>::p3* p3, int64_t n,
               const FT* const* in, FT* const* out, double* diag, void* stream) {
    if (!p1 || !p2 || !p3) return cmh::fail(CUMICRO_E_NULL, "fused: a parameter block is NULL");
    if (!in || !out) return cmh::fail(CUMICRO_E_NULL, "fused: column pointer table is NULL");
    if (n < 0) return cmh::fail(CUMICRO_E_SIZE, "n = %lld is negative", (long long)n);
    for (int c = 0; c < NIN; ++c)
        if (n > 0 && in[c] == nullptr) return cmh::fail(CUMICRO_E_NULL, "fused: input column %d is NULL", c);
    if (p3->n_modes < 0 || p3->n_modes > kMaxModes) return cmh::fail(CUMICRO_E_OPTION, "n_modes = %d", (int)p3->n_modes);
    cudaStream_t s = (cudaStream_t)stream;
    FusedArgs<FT> a;
    widen(*p1, a.f.p1);
    widen(*p2, a.f.p2);
    widen(*p3, a.f.p3);
    a.f.tk = make_thermo_k<D>(a.f.p1.tps, is_f32<FT>());
    a.f.k1 = make_1m_k<D>(a.f.p1, is_f32<FT>());
    a.f.k2 = make_sb2006_k<D>(a.f.p2.sb, a.f.p2.aps, is_f32<FT>());
    a.f.k3 = make_arg_k<D>(a.f.p3, is_f32<FT>());
    a.f.with_activation = a.f.p3.n_modes > 0;
    for (int c = 0; c < NIN; ++c) a.in[c] = in[c];
    for (int c = 0; c < NOUT; ++c) a.out[c] = out[c];
    a.n = n;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + BLOCK - 1) / BLOCK, (int64_t)cmh::num_sms() * 4 * 4));
    void* ws = nullptr;
    int st = cmh::workspace(cmh::kPipeSlots /* slot reserved for the diagnostics partials */, sizeof(double) * NDIAG * (size_t)blocks + 64, &ws);
    if (st) return st;
    a.partials = static_cast<double*>(ws);
    fused_kernel<FT><<<blocks, BLOCK, 0, s>>>(a);
    cmh::count_launch();
    if (diag) {
        fused_diag_finish<<<1, 32, 0, s>>>(a.partials, blocks, diag);
        cmh::count_launch();
    }
    return cmh::cuda_status(cudaGetLastError(), "fused_1m2m_icenuc launch");
}

}  // namespace

extern "C" {

int cumicro_fused_1m2m_icenuc_f64(const cumicro_params_1m_f64* p1, const cumicro_params_2m_warm_f64* p2,
                                  const cumicro_params_icenuc_f64* p3, int64_t n, const double* const* in11,
                                  double* const* out11, double* diag, void* stream) {
    return fused_impl<double>(p1, p2, p3, n, in11, out11, diag, stream);
}
int cumicro_fused_1m2m_icenuc_f32(const cumicro_params_1m_f32* p1, const cumicro_params_2m_warm_f32* p2,
                                  const cumicro_params_icenuc_f32* p3, int64_t n, const float* const* in11,
                                  float* const* out11, double* diag, void* stream) {
    return fused_impl<float>(p1, p2, p3, n, in11, out11, diag, stream);
}

}  // extern "C"
