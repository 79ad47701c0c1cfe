// kernels_fused.cu — the "ClimaAtmos-scale" fused kernel of BASELINE config 5:
// 1-moment tendencies + 2-moment warm-rain tendencies + ice-nucleation rates (+ ARG2000
// activated number) from ONE read of the 11 state columns, all 11 outputs written once,
// and the domain diagnostics (precipitation production, activated number) reduced in the
// kernel epilogue: warp shuffles -> one partial per block -> a fixed-order second pass, so the
// sums are bit-reproducible and no output column is re-read.  The cross-GPU sum of the
// per-GPU diagnostics is the only collective of the whole path (NCCL all-reduce of
// CUMICRO_NDIAG doubles, done by the host binding).
#include <cmath>
#include <cstdlib>
#include <limits>
#include <string>

#include "cm_1m.cuh"
#include "cm_hostpipe.cuh"
#include "cm_icenuc.cuh"
#include "cm_launch.cuh"
#include "cm_p2p.cuh"
#include "cm_sb2006.cuh"
#include "cm_sb2006_fast.cuh"
#include "cm_tile2m.cuh"

namespace {

using namespace cm;
using D = double;
template <class FT> constexpr bool is_f32() { return sizeof(FT) == 4; }

constexpr int NIN = 11;   // rho, T, p, w, q_tot, q_lcl, q_icl, q_rai, q_sno, n_lcl, n_rai
constexpr int NOUT = 11;  // 1M: dq_lcl, dq_icl, dq_rai, dq_sno | 2M: dq_lcl, dn_lcl, dq_rai, dn_rai | J_dep, J_ABIFM, J_hom
constexpr int NDIAG = CUMICRO_NDIAG;
// Launch shape.  The loop body is the code of three kernel families (~80 KB of SASS) against a 32 KB L1.5 instruction cache:
// with independent 128-thread blocks the resident warps sit in different families and thrash it (ncu: stall_no_instruction 1.9
// per issue vs 0.15 in the 2M kernel alone).  ONE block of 768 threads per SM (optionally a block barrier after every family) keeps
// all 24 warps of the SM inside the same family's code.   Measured, 2^24 points (tools/tune_fused.py; n = no barriers):
//   128x6n 3.27   128x6 3.17   256x3n 3.12   256x3 2.88   384x2n 2.98   384x2 2.82   768x1n 2.77   768x1 2.93 ms
// -> one block per SM WITHOUT barriers: its warps start together and stay loosely in phase; the barrier's wait for the
// slowest warp then costs more than the residual drift.
#ifndef CUMICRO_FUSED_BLOCK
#define CUMICRO_FUSED_BLOCK 896   /* one block per SM: 640 2.53, 768 2.47, 896 2.42, 1024 2.47 ms (after the SB2006 specialisation) */
#define CUMICRO_FUSED_MINB 1
#endif

struct FusedParams {
    P<D>::params_1m p1;
    P<D>::params_2m_warm p2;
    cumicro_params_icenuc_f64 p3;
    ThermoK<D> tk;
    OneMK<D> k1;
    SB2006K<D> k2;
    W2K w2k;            // constants of the fast 2-moment body (SPEC >= 0)
    ArgK<D> k3;
    int with_activation;
};

template <class FT> struct FusedArgs {
    FusedParams f;
    const FT* in[NIN];
    FT* out[NOUT];
    double* partials;   // [gridDim.x][NDIAG]
    const double* tab;  // ventilation table of the 2-moment block (cmh::w2_table) or nullptr
    int64_t n;
};

// The three families run one after the other on the same point.  Their inputs wait in shared memory (per-thread cp.async
// copies, double-buffered: the next item streams in while this one is computed) and every family stores its tendencies and
// accumulates its diagnostic as soon as it is done, so the live register set is that of ONE family at a time and the kernel
// keeps the occupancy of the single-family kernels.
// S1M: the 1-moment block has the default exponent structure (cm_1m.cuh, OneMK::std_exponents): powers of λ⁻¹ by multiplication
// ALL_OUT: every output column is wanted (decided at launch): no per-column NULL test in the store sequence
// NM3: the aerosol distribution has exactly three modes (compile-time trip count of the ARG2000 mode loops)
template <class FT, int BLOCK, int MINB, bool SYNC, int SPEC, bool TAB, bool S1M, bool ALL_OUT, bool NM3>
__global__ void __launch_bounds__(BLOCK, MINB) fused_kernel(const __grid_constant__ FusedArgs<FT> a) {
    __shared__ __align__(16) double tab_s[TAB ? kTabDoubles : 2];
    if (TAB) {
        for (int i = threadIdx.x; i < kTabDoubles / 2; i += BLOCK)
            reinterpret_cast<double2*>(tab_s)[i] = __ldg(reinterpret_cast<const double2*>(a.tab) + i);
    }
    math_tables_init<BLOCK>();
    if constexpr (SPEC >= 0) math_tables_init_log2<BLOCK>();
    extern __shared__ __align__(16) unsigned char fused_dyn_smem[];
    FT (*stage)[NIN][BLOCK] = reinterpret_cast<FT (*)[NIN][BLOCK]>(fused_dyn_smem);   // [2][NIN][BLOCK]
    const FusedParams& f = a.f;
    const int tid = threadIdx.x;
    double diag[NDIAG];
#pragma unroll
    for (int k = 0; k < NDIAG; ++k) diag[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * BLOCK;
    int64_t base = (int64_t)blockIdx.x * BLOCK;   // block-uniform trip count: the barriers sit inside the loop
    if (base + tid < a.n) {
#pragma unroll
        for (int c = 0; c < NIN; ++c) cp_async_elem(&stage[0][c][tid], a.in[c] + base + tid);
    }
    cp_async_commit();
    int buf = 0;
    for (; base < a.n; base += stride) {
        const int64_t i = base + tid;
        const bool active = i < a.n;
        const int64_t nxt = i + stride;
        if (nxt < a.n) {
#pragma unroll
            for (int c = 0; c < NIN; ++c) cp_async_elem(&stage[buf ^ 1][c][tid], a.in[c] + nxt);
        }
        cp_async_commit();
        cp_async_wait<1>();
        auto in = [&](int c) { return (double)stage[buf][c][tid]; };   // rho, T, p, w, q_tot, q_lcl, q_icl, q_rai, q_sno, n_lcl, n_rai
        auto put = [&](int c, double v) { if (ALL_OUT || a.out[c]) __stcs(a.out[c] + i, (FT)v); };
        // 1-moment tendencies                                         BMT:505-514
        // the temperature-only thermodynamic state, ONCE for the 1-moment and the ice-nucleation / ARG2000 families, which run
        // back to back (the 2-moment body has its own log-space form and runs last, so nothing is held across it)
        ThermoShared<D> th;
        if (active) th = thermo_shared<D>(f.tk, in(1));
        if (active) {
            const Src1M<D> r = microphysics_source_terms_1m<D, S1M>(f.p1, f.tk, f.k1, in(0), in(1), in(4), in(5), in(6), in(7), in(8), &th);
            double t[4];
            aggregate_tendencies_1m<D>(r, t);
#pragma unroll
            for (int k = 0; k < 4; ++k) put(k, t[k]);
            diag[0] += in(0) * (t[2] + t[3]);   // 1M precipitation production  Σ ρ (dq_rai + dq_sno)   [kg m^-3 s^-1]
        }
        if (SYNC) __syncthreads(); else asm volatile("" ::: "memory");
        // ice-nucleation rates (+ ARG2000 activated number)              IN:92-134, 557-584; AA:138-273
        if (active) {
            const double rho = in(0), T = in(1), pr = in(2), w = in(3), q_tot = in(4), q_lcl = in(5), q_icl = in(6), q_rai = in(7),
                         q_sno = in(8), n_lcl = in(9);
            double n_act = 0.0;
            double da_w;
            if (f.with_activation) {
                const ArgOut o = arg2000<false, false, (NM3 ? 3 : -1)>(f.p3, f.tk, f.k3, T, pr, w, q_tot, q_lcl + q_rai, q_icl + q_sno, rho * n_lcl, 0.0, &th);
                da_w = o.da_w;
                // Activation needs an updraft: AA.max_supersaturation takes sqrt(alpha w / G) (AA:170-176), a DomainError in the
                // reference for w < 0 and S_max = 0 for w = 0.  A model slab has downdraft cells: they activate nothing, and they
                // must not poison the domain sum (NaN from one cell would make the all-reduced diagnostic NaN everywhere).
                const bool updraft = w > 0.0;
#pragma unroll
                for (int m = 0; m < (NM3 ? 3 : kMaxModes); ++m)
                    if (NM3 || m < f.p3.n_modes) {
                        const double na = o.N_act[m];
                        n_act += (updraft && na == na && na < 1.7e308) ? na : 0.0;
                    }
            } else {
                const D pl = th.p_vs_l;
                const D Rm = f.p3.tps.R_d * (1.0 + (f.k3.Rv_over_Rd - 1.0) * q_tot - f.k3.Rv_over_Rd * (q_lcl + q_rai + q_icl + q_sno));
                const D p_v = (q_tot - (q_lcl + q_rai) - (q_icl + q_sno)) * (pr / (Rm * T)) * f.tk.R_v * T;
                da_w = p_v / pl - th.p_vs_i / pl;
            }
            bool err = false;
            put(8, deposition_J<D>(f.p3.dust, da_w, f.k3.ln10));
            put(9, ABIFM_J<D>(f.p3.dust, da_w, f.k3.ln10));
            double jh = f.p3.hom_linear ? homogeneous_J_linear<D>(f.p3.koop, da_w, f.k3.ln10)
                                        : homogeneous_J_cubic<D>(f.p3.koop, da_w, f.k3.ln10, err);
            if (err) jh = __longlong_as_double(0x7ff8000000000000LL);
            put(10, jh);
            diag[2] += n_act;                   // activated aerosol number     Σ N_act                  [m^-3]
            diag[3] += 1.0;                     // points
        }
        if (SYNC) __syncthreads(); else asm volatile("" ::: "memory");
        // 2-moment warm rain (cloud ice seen by the thermodynamics = q_icl + q_sno)   BMT:820-854
        if (active) {
            const double q_ice = clamp0_(in(6)) + clamp0_(in(8));
            if constexpr (SPEC >= 0) {   // default SB2006 block structure: the headline body (cm_sb2006_fast.cuh), bit-identical to cumicro_bmt2m_warm_*
                double y[4];
                warm2m_fast<SPEC, TAB>(f.w2k, in(0), in(1), in(4), in(5), in(9), in(7), in(10), q_ice, true, y, tab_s);
#pragma unroll
                for (int k = 0; k < 4; ++k) put(4 + k, y[k]);
                diag[1] += in(0) * y[2];        // 2M rain production           Σ ρ dq_rai
            } else {
                const Warm2M<D> o = warm_rain_tendencies_2m<D, -1>(f.p2, f.tk, f.k2, in(0), in(1), in(4), in(5), in(9), in(7), in(10), q_ice);
                put(4, o.dq_lcl_dt);
                put(5, o.dn_lcl_dt);
                put(6, o.dq_rai_dt);
                put(7, o.dn_rai_dt);
                diag[1] += in(0) * o.dq_rai_dt;
            }
        }
        if (SYNC) __syncthreads(); else asm volatile("" ::: "memory");
        buf ^= 1;
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

// fixed-order second pass: one block; thread t sums the partials of blocks t, t + 256, ... in that order, then a fixed
// binary tree over the 256 thread sums -> bit-reproducible for a given grid size
// P2P: the cross-GPU sum is the tail of this kernel (peer-memory stores over NVLink, cm_p2p.cuh) — diag leaves it as the DOMAIN sums.
template <bool P2P>
__global__ void __launch_bounds__(256) fused_diag_finish(const double* partials, int n_blocks, double* diag, const __grid_constant__ cm::P2PDev p2p) {
    __shared__ double sh[256];
    for (int k = 0; k < NDIAG; ++k) {
        double v = 0.0;
        for (int b = threadIdx.x; b < n_blocks; b += 256) v += partials[(size_t)b * NDIAG + k];
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
            __syncthreads();
        }
        if (threadIdx.x == 0) diag[k] = sh[0];
        __syncthreads();
    }
    if (P2P) cm::p2p_allreduce_block(p2p, diag, NDIAG);
}

template <class FT> struct PF;
template <> struct PF<double> { using p1 = cumicro_params_1m_f64; using p2 = cumicro_params_2m_warm_f64; using p3 = cumicro_params_icenuc_f64; };
template <> struct PF<float> { using p1 = cumicro_params_1m_f32; using p2 = cumicro_params_2m_warm_f32; using p3 = cumicro_params_icenuc_f32; };

template <class FT, int BLOCK, int MINB, bool SYNC, int SPEC, bool TAB = false, bool S1M = false, bool ALL_OUT = false, bool NM3 = false> int launch_fused(FusedArgs<FT>& a, int64_t n, cudaStream_t s, double* diag, const cm::P2PDev* p2p) {
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + BLOCK - 1) / BLOCK, (int64_t)cmh::num_sms() * MINB * 16));   // 16 waves of the resident grid (cm_launch.cuh)
    // the block partials: scratch of this (device, stream) — calls in flight on different streams never share it, and work on ONE
    // stream is ordered (the finish kernel of call k has read the partials before the main kernel of call k + 1 writes them)
    void* ws = nullptr;
    int st = cmh::stream_workspace(s, sizeof(double) * NDIAG * (size_t)blocks + 64, &ws);
    if (st) return st;
    a.partials = static_cast<double*>(ws);
    const size_t smem = sizeof(FT) * 2 * NIN * BLOCK;
    auto kern = fused_kernel<FT, BLOCK, MINB, SYNC, SPEC, TAB, S1M, ALL_OUT, NM3>;
    {   // every launch: function attributes are per device, and the call costs about a microsecond against a millisecond kernel
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cmh::cuda_status(e, "fused: cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    }
    kern<<<blocks, BLOCK, smem, s>>>(a);
    cmh::count_launch();
    if (diag) {
        if (p2p) fused_diag_finish<true><<<1, 256, 0, s>>>(a.partials, blocks, diag, *p2p);
        else fused_diag_finish<false><<<1, 256, 0, s>>>(a.partials, blocks, diag, cm::P2PDev{});
        cmh::count_launch();
    }
    return CUMICRO_OK;
}

template <class FT>
int fused_impl(const typename PF<FT>::p1* p1, const typename PF<FT>::p2* p2, const typename PF<FT>::p3* p3, int64_t n,
               const FT* const* in, FT* const* out, double* diag, void* stream, void* win = nullptr, bool want_p2p = false) {
    cm::P2PDev p2p_dev{};
    const cm::P2PDev* p2p = nullptr;
    if (want_p2p) {   // the exchange rides on the finish kernel: it needs the diagnostics vector
        if (diag == nullptr) return cmh::fail(CUMICRO_E_NULL, "fused (p2p): diag is NULL");
        const int stw = cmh::p2p_dev(win, &p2p_dev);
        if (stw) return stw;
        p2p = &p2p_dev;
    }
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
    a.f.w2k = make_w2k(a.f.p2, is_f32<FT>());
    a.f.k3 = make_arg_k<D>(a.f.p3, is_f32<FT>());
    a.f.with_activation = a.f.p3.n_modes > 0;
    for (int c = 0; c < NIN; ++c) a.in[c] = in[c];
    for (int c = 0; c < NOUT; ++c) a.out[c] = out[c];
    a.n = n;
    // launch shape: one block of CUMICRO_FUSED_BLOCK threads per SM, no family barriers (see the top of the file; the sweep was run with every shape instantiated).
    // CUMICRO_FUSED_SHAPE=128x6 selects the old shape for comparison.  The 2-moment family runs the instantiation specialised
    // for the default SB2006 block structure when the block has it (cm_sb2006.cuh, sb2006_spec()).
    bool small_blocks = false;
    if (const char* e = std::getenv("CUMICRO_FUSED_SHAPE")) small_blocks = std::string(e).rfind("128", 0) == 0;
    const int spec = w2k_supported(a.f.p2) ? (a.f.p2.sb.pdf_r.limited ? 1 : 0) : -1;
    a.tab = (spec == 1) ? cmh::w2_table(a.f.p2, a.f.w2k) : nullptr;
    bool all_out = true;
    for (int c = 0; c < NOUT; ++c) all_out = all_out && out[c] != nullptr;
    int st;
    if (small_blocks) st = launch_fused<FT, 128, 6, false, -1>(a, n, s, diag, p2p);
    else if (spec == 1 && a.tab && a.f.k1.std_exponents && all_out && a.f.p3.n_modes == 3)
        st = launch_fused<FT, CUMICRO_FUSED_BLOCK, 1, false, 1, true, true, true, true>(a, n, s, diag, p2p);
    else if (spec == 1 && a.tab && a.f.k1.std_exponents && all_out) st = launch_fused<FT, CUMICRO_FUSED_BLOCK, 1, false, 1, true, true, true>(a, n, s, diag, p2p);
    else if (spec == 1 && a.tab && a.f.k1.std_exponents) st = launch_fused<FT, CUMICRO_FUSED_BLOCK, 1, false, 1, true, true>(a, n, s, diag, p2p);
    else if (spec == 1 && a.tab) st = launch_fused<FT, CUMICRO_FUSED_BLOCK, 1, false, 1, true>(a, n, s, diag, p2p);
    else if (spec == 1) st = launch_fused<FT, CUMICRO_FUSED_BLOCK, 1, false, 1>(a, n, s, diag, p2p);
    else if (spec == 0) st = launch_fused<FT, CUMICRO_FUSED_BLOCK, 1, false, 0>(a, n, s, diag, p2p);
    else st = launch_fused<FT, CUMICRO_FUSED_BLOCK, 1, false, -1>(a, n, s, diag, p2p);
    if (st) return st;
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
int cumicro_fused_1m2m_icenuc_p2p_f64(const cumicro_params_1m_f64* p1, const cumicro_params_2m_warm_f64* p2,
                                      const cumicro_params_icenuc_f64* p3, int64_t n, const double* const* in11,
                                      double* const* out11, double* diag, void* win, void* stream) {
    return fused_impl<double>(p1, p2, p3, n, in11, out11, diag, stream, win, true);
}
int cumicro_fused_1m2m_icenuc_p2p_f32(const cumicro_params_1m_f32* p1, const cumicro_params_2m_warm_f32* p2,
                                      const cumicro_params_icenuc_f32* p3, int64_t n, const float* const* in11,
                                      float* const* out11, double* diag, void* win, void* stream) {
    return fused_impl<float>(p1, p2, p3, n, in11, out11, diag, stream, win, true);
}

}  // extern "C"
