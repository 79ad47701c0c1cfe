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
#include "cm_sb2006_fast.cuh"
#include "cm_tile2m.cuh"

#include <cstring>
#include <mutex>
#include <vector>

namespace cmh {
// Ventilation tables: one per (device, parameter-block content), built and verified on the host, uploaded once.
// launch shapes of the generic (non-default structure) 2-moment body and of the 15-column leaves kernel: tile shape or grid-stride
#ifndef CUMICRO_2MG_TILED
#define CUMICRO_2MG_TILED 0   /* generic body, 2^24 points: 0.627 ms pipelined grid-stride, 0.622-0.626 tiles: no difference */
#endif
#ifndef CUMICRO_2MG_MINB
#define CUMICRO_2MG_MINB 6
#endif
#ifndef CUMICRO_2ML_TILED
#define CUMICRO_2ML_TILED 1   /* 15 leaf columns, 2^24 points: 1.111 ms grid-stride (128x4) -> 0.622 ms tiles (128x5; x4 0.632, x6 0.627) */
#endif
#ifndef CUMICRO_2ML_MINB
#define CUMICRO_2ML_MINB 5
#endif
namespace {
struct TabEntry {
    int device;
    cumicro_params_2m_warm_f64 key;
    double eps_n, tab_inv_h, tab_u0;
    double* dev;      // nullptr: verification failed for this block (closed form is used)
};
std::mutex g_tab_mutex;
std::vector<TabEntry> g_tabs;
}  // namespace

const double* w2_table(const cumicro_params_2m_warm_f64& p, cm::W2K& k) {
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    for (const TabEntry& e : g_tabs)
        if (e.device == device && e.eps_n == k.eps_n && std::memcmp(&e.key, &p, sizeof(p)) == 0) {
            k.tab_inv_h = e.tab_inv_h; k.tab_u0 = e.tab_u0;
            return e.dev;
        }
    std::vector<double> host(cm::kTabDoubles);
    TabEntry e{};
    e.device = device; std::memcpy(&e.key, &p, sizeof(p)); e.eps_n = k.eps_n; e.dev = nullptr;
    const double err = cm::build_w2_table(p, k, host.data());
    e.tab_inv_h = k.tab_inv_h; e.tab_u0 = k.tab_u0;
    if (err < 1e-15) {
        double* d = nullptr;
        // not cached on failure (e.g. a stream capture in progress forbids the allocation): the next call tries again
        if (cudaMalloc(&d, sizeof(double) * cm::kTabDoubles) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (cudaMemcpy(d, host.data(), sizeof(double) * cm::kTabDoubles, cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError(); cudaFree(d); return nullptr;
        }
        e.dev = d;
    }
    g_tabs.push_back(e);
    return e.dev;
}
void release_tables() {
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    for (TabEntry& e : g_tabs)
        if (e.dev) cudaFree(e.dev);
    g_tabs.clear();
}
}  // namespace cmh

namespace {

using namespace cm;

// ---- BMT:820-854: 7 columns in, 4 tendencies out --------------------------------------
// (NIN = 8 adds the optional q_ice column that the reference's warm-only method
// accepts and forwards to the thermodynamics, BMT:823,836,843.)
// Functors compute in Float64 (D = double); the entry points below are templated on the
// column type FT and widen Float32 parameter blocks exactly.
using D = double;
template <class FT> constexpr bool is_f32() { return sizeof(FT) == 4; }
// SPEC: compile-time specialisation for the default structure of the SB2006 block (cm_sb2006.cuh, sb2006_spec()); -1 = generic.
template <int NIN = 7, int SPEC = -1> struct Warm2MFused {
    P<D>::params_2m_warm p;
    ThermoK<D> tk;
    SB2006K<D> sk;
    __device__ __forceinline__ void operator()(const D (&x)[NIN], D (&y)[4]) const {
        const D q_ice = (NIN == 8) ? clamp0_(x[NIN - 1]) : D(0);
        Warm2M<D> o = warm_rain_tendencies_2m<D, SPEC>(p, tk, sk, x[0], x[1], x[2], x[3], x[4], x[5], x[6], q_ice);
        y[0] = o.dq_lcl_dt;
        y[1] = o.dn_lcl_dt;
        y[2] = o.dq_rai_dt;
        y[3] = o.dn_rai_dt;
    }
};

// The headline body (cm_sb2006_fast.cuh): default SB2006 block structure, LIM = limited rain PSD.
template <int NIN, int LIM> struct Warm2MFast {
    static constexpr bool kNeedsLog2 = true;
    W2K k;
    __device__ __forceinline__ void operator()(const D (&x)[NIN], D (&y)[4]) const {
        warm2m_fast<LIM>(k, x[0], x[1], x[2], x[3], x[4], x[5], x[6], (NIN == 8) ? clamp0_(x[NIN - 1]) : D(0), NIN == 8, y);
    }
};
template <class FT, class F> F make_2m_fast(const typename P<FT>::params_2m_warm* p) {
    F f{};
    P<D>::params_2m_warm w;
    widen(*p, w);
    f.k = make_w2k(w, is_f32<FT>());
    return f;
}
template <class FT> const double* w2_table_for(const typename P<FT>::params_2m_warm* p, W2K& k) {
    P<D>::params_2m_warm w;
    widen(*p, w);
    return cmh::w2_table(w, k);
}
template <class FT> bool fast_ok(const typename P<FT>::params_2m_warm* p) {
    P<D>::params_2m_warm w;
    widen(*p, w);
    return w2k_supported(w);
}

// ---- the 15 SB2006 process rates one by one (leaf API) ----------------------------------
struct Warm2MLeaves {
    P<D>::params_2m_warm p;
    ThermoK<D> tk;
    SB2006K<D> sk;
    __device__ __forceinline__ void operator()(const D (&x)[7], D (&y)[CUMICRO_SB2006_NLEAF]) const {
        Warm2M<D> o = warm_rain_tendencies_2m<D>(p, tk, sk, x[0], x[1], x[2], x[3], x[4], x[5], x[6], D(0));
#pragma unroll
        for (int k = 0; k < CUMICRO_SB2006_NLEAF; ++k) y[k] = o.leaf[k];
    }
};

// (p wide, tk, sk) of one call: Float32 blocks are widened exactly, thresholds follow FT.
// CM2.rain_evaporation(sb, aps, tps, q_tot, q_lcl, q_icl, q_rai, q_sno, rho, N_rai, T) (CM2:780-828) and
// CM2.∂rain_evaporation_∂N_rai_∂q_rai (CM2:844-853): columns q_tot, q_lcl, q_icl, q_rai, q_sno, rho, N_rai, T -> 4 columns
struct RainEvap2M {
    P<D>::params_2m_warm p;
    ThermoK<D> tk;
    SB2006K<D> sk;
    __device__ __forceinline__ void operator()(const D (&x)[8], D (&y)[4]) const {
        const D q_tot = x[0], q_lcl = x[1], q_icl = x[2], q_rai = x[3], q_sno = x[4], rho = x[5], N_rai = x[6], T = x[7];
        // q_liq = q_lcl + q_rai, q_ice = q_icl + q_sno (CM2:784); negative inputs are clamped to 0 as in the BMT caller (BMT:827-836)
        const Warm2M<D> o = warm_rain_tendencies_2m<D>(p, tk, sk, rho, T, q_tot, q_lcl, D(0), q_rai, D(0), q_icl + q_sno, fmax_(N_rai, D(0)));
        const D dn = o.leaf[CUMICRO_SB_EVAP_DN_RAI], dq = o.leaf[CUMICRO_SB_EVAP_DQ_RAI];
        y[0] = dn;
        y[1] = dq;
        y[2] = (N_rai > tk.eps) ? div_(dn, N_rai) : D(0);     // ∂(∂ₜρn_rai/ρ)/∂N_rai ≈ ∂ₜρn_rai / N_rai
        y[3] = (q_rai > tk.eps) ? div_(dq, q_rai) : D(0);     // ∂(∂ₜq_rai)/∂q_rai ≈ ∂ₜq_rai / q_rai
    }
};

template <class FT, class F> F make_2m(const typename P<FT>::params_2m_warm* p) {
    F f{};
    widen(*p, f.p);
    f.tk = make_thermo_k<D>(f.p.tps, is_f32<FT>());
    f.sk = make_sb2006_k<D>(f.p.sb, f.p.aps, is_f32<FT>());
    return f;
}

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
    const char* w = "bmt2m_warm kernel launch";
    const bool fast = fast_ok<FT>(p);
    const bool lim = p->sb.pdf_r.limited != 0;
    if (q_ice != nullptr) {
        const FT* in8[8] = {rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice};
        if (fast) {
            W2K k = make_2m_fast<FT, Warm2MFast<8, 1>>(p).k;
            const double* tab = lim ? w2_table_for<FT>(p, k) : nullptr;
            return lim ? launch_warm2m_tile<FT, 8, 1>(k, tab, n, in8, out, s, w) : launch_warm2m_tile<FT, 8, 0>(k, nullptr, n, in8, out, s, w);
        }
        return launch_pointwise<FT, 8, 4, Warm2MFused<8>, 128, 8, false>(make_2m<FT, Warm2MFused<8>>(p), n, in8, out, s, w);
    }
#ifdef CUMICRO_TUNING
    if (fast && lim) {   // launch-shape exploration (tools/tune_2m.py); not compiled into the product build
        const char* ev = getenv("CUMICRO_2M_VARIANT");
        const int v = ev ? atoi(ev) : 0;
        using F1 = Warm2MFast<7, 1>;
        const F1 f1 = make_2m_fast<FT, F1>(p);
        W2K kv = f1.k;
        const char* nt = getenv("CUMICRO_NO_TABLE");
        const double* tabv = (nt && nt[0] == '1') ? nullptr : w2_table_for<FT>(p, kv);
        switch (v) {
            case 13: return launch_pointwise<FT, 7, 4, Warm2MFused<7, 1>, 128, 7, false, true>(make_2m<FT, Warm2MFused<7, 1>>(p), n, in, out, s, w);  // round-1 kernel
            case 14: return launch_pointwise<FT, 7, 4, F1, 128, 7, false, true>(f1, n, in, out, s, w);   // fast body, cp.async shape
            case 20: return launch_warm2m_tile<FT, 7, 1, 128, 7, 1>(kv, tabv, n, in, out, s, w);
            case 21: return launch_warm2m_tile<FT, 7, 1, 128, 8, 1>(kv, tabv, n, in, out, s, w);
            case 22: return launch_warm2m_tile<FT, 7, 1, 128, 6, 1>(kv, tabv, n, in, out, s, w);
            case 23: return launch_warm2m_tile<FT, 7, 1, 256, 3, 1>(kv, tabv, n, in, out, s, w);
            case 24: return launch_warm2m_tile<FT, 7, 1, 128, 5, 1>(kv, tabv, n, in, out, s, w);
            case 25: return launch_warm2m_tile<FT, 7, 1, 64, 12, 1>(kv, tabv, n, in, out, s, w);
            case 26: return launch_warm2m_tile<FT, 7, 1, 192, 4, 1>(kv, tabv, n, in, out, s, w);
            // two points per thread share every constant load and the tile bookkeeping (1456 SASS instructions for two points against
            // 864 for one, no spills at 96 registers) and are still slower — the body is not bound by instruction issue alone
            // (tools/tune_2m_ab.py, round-robin, ms per 2^24 points): 128x6 PPT 1 0.398 | PPT 2: 128x3 0.452, 128x4 0.428, 128x5 0.414, 96x6 0.439
            case 40: return launch_warm2m_tile<FT, 7, 1, 128, 3, 2>(kv, tabv, n, in, out, s, w);
            case 41: return launch_warm2m_tile<FT, 7, 1, 128, 4, 2>(kv, tabv, n, in, out, s, w);
            case 42: return launch_warm2m_tile<FT, 7, 1, 64, 6, 2>(kv, tabv, n, in, out, s, w);
            case 43: return launch_warm2m_tile<FT, 7, 1, 64, 8, 2>(kv, tabv, n, in, out, s, w);
            case 44: return launch_warm2m_tile<FT, 7, 1, 128, 5, 2>(kv, tabv, n, in, out, s, w);
            case 47: return launch_warm2m_tile<FT, 7, 1, 96, 6, 2>(kv, tabv, n, in, out, s, w);
            // other block sizes at one point per thread (same tool): 128x6 0.398 | 128x7 0.402, 128x8 0.411, 96x8 0.416, 96x9 0.419, 160x5 0.432, 224x4 0.436, 160x4 0.439
            case 50: return launch_warm2m_tile<FT, 7, 1, 96, 8, 1>(kv, tabv, n, in, out, s, w);
            case 51: return launch_warm2m_tile<FT, 7, 1, 160, 5, 1>(kv, tabv, n, in, out, s, w);
            case 52: return launch_warm2m_tile<FT, 7, 1, 224, 4, 1>(kv, tabv, n, in, out, s, w);
            case 53: return launch_warm2m_tile<FT, 7, 1, 96, 9, 1>(kv, tabv, n, in, out, s, w);
            case 54: return launch_warm2m_tile<FT, 7, 1, 160, 4, 1>(kv, tabv, n, in, out, s, w);
            default: break;
        }
    }
#endif
    // 128 x 7 blocks/SM: sweep in tools/tune_2m.py.  The default parameter STRUCTURE (exponents 3 / 4 / -5, any values) runs the
    // fast body (tile kernel, cm_tile2m.cuh; cp.async shape for columns that are not 16-byte aligned); anything else the generic one.
    if (fast) {
        W2K k = make_2m_fast<FT, Warm2MFast<7, 1>>(p).k;
        const double* tab = lim ? w2_table_for<FT>(p, k) : nullptr;
        return lim ? launch_warm2m_tile<FT, 7, 1>(k, tab, n, in, out, s, w) : launch_warm2m_tile<FT, 7, 0>(k, nullptr, n, in, out, s, w);
    }
#if CUMICRO_2MG_TILED
    return launch_pointwise_tiled<FT, 7, 4, Warm2MFused<7>, 128, CUMICRO_2MG_MINB>(make_2m<FT, Warm2MFused<7>>(p), n, in, out, s, w);
#else
    return launch_pointwise<FT, 7, 4, Warm2MFused<7>, 128, 7, false, true>(make_2m<FT, Warm2MFused<7>>(p), n, in, out, s, w);
#endif
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
#if CUMICRO_2ML_TILED
    return launch_pointwise_tiled<FT, 7, CUMICRO_SB2006_NLEAF, Warm2MLeaves, 128, CUMICRO_2ML_MINB>(make_2m<FT, Warm2MLeaves>(p), n, in, out,
                                                                                                  (cudaStream_t)stream, "sb2006_leaves launch");
#endif
    return launch_pointwise<FT, 7, CUMICRO_SB2006_NLEAF, Warm2MLeaves, 128, 4, false>(make_2m<FT, Warm2MLeaves>(p), n, in, out,
                                                                                     (cudaStream_t)stream, "sb2006_leaves kernel launch");
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
    const bool fast = fast_ok<FT>(p);
    const bool lim = p->sb.pdf_r.limited != 0;
    const Warm2MFused<7> f = make_2m<FT, Warm2MFused<7>>(p);
    const Warm2MFast<7, 1> f1 = make_2m_fast<FT, Warm2MFast<7, 1>>(p);
    W2K kt = f1.k;
    const double* tab = (fast && lim) ? w2_table_for<FT>(p, kt) : nullptr;
    return host_pipeline<FT, 7, 4>(n, in, out, chunk,
                                   [&](int64_t m, const FT* const(&din)[7], FT* const(&dout)[4], cudaStream_t s) {
                                       const char* w = "bmt2m_warm (host pipeline) kernel launch";
                                       if (fast)
                                           return lim ? launch_warm2m_tile<FT, 7, 1>(kt, tab, m, din, dout, s, w) : launch_warm2m_tile<FT, 7, 0>(kt, nullptr, m, din, dout, s, w);
                                       return launch_pointwise<FT, 7, 4, Warm2MFused<7>, 128, 7, false, true>(f, m, din, dout, s, w);
                                   });
}

template <class FT>
int rain_evaporation_impl(const typename P<FT>::params_2m_warm* p, int64_t n, const FT* const* in8, FT* const* out4, void* stream) {
    if (in8 == nullptr || out4 == nullptr) return cmh::fail(CUMICRO_E_NULL, "column pointer table is NULL");
    const FT* in[8];
    FT* out[4];
    for (int c = 0; c < 8; ++c) in[c] = in8[c];
    for (int c = 0; c < 4; ++c) out[c] = out4[c];
    int st = validate_columns<FT, 8>(p, n, in);
    if (st) return st;
    if ((st = check_2m_options<FT>(p))) return st;
#if CUMICRO_2ML_TILED
    return launch_pointwise_tiled<FT, 8, 4, RainEvap2M, 128, CUMICRO_2ML_MINB>(make_2m<FT, RainEvap2M>(p), n, in, out, (cudaStream_t)stream,
                                                                 "rain_evaporation_2m kernel launch");
#else
    return launch_pointwise<FT, 8, 4, RainEvap2M, 128, 4, false>(make_2m<FT, RainEvap2M>(p), n, in, out, (cudaStream_t)stream,
                                                                 "rain_evaporation_2m kernel launch");
#endif
}

// ---- terminal velocities: (q, rho, N) -> (vt0, vt1) ---------------------------------------
struct RainVelSB {
    P<D>::sb_pdf_r pdf_r;
    P<D>::vel_sb2006 vel;
    D pi_rho_w, eps;
    __device__ __forceinline__ void operator()(const D (&x)[3], D (&y)[2]) const {
        rain_terminal_velocity_sb<D>(pdf_r, vel, pi_rho_w, eps, x[0], x[1], x[2], y[0], y[1]);
    }
};
struct RainVelChen {
    P<D>::sb_pdf_r pdf_r;
    P<D>::vel_chen_rain vel;
    D pi_rho_w, eps;
    __device__ __forceinline__ void operator()(const D (&x)[3], D (&y)[2]) const {
        rain_terminal_velocity_chen<D>(pdf_r, vel, pi_rho_w, eps, x[0], x[1], x[2], y[0], y[1]);
    }
};
struct CloudVel {
    P<D>::sb_pdf_c pdf_c;
    P<D>::vel_stokes vel;
    D pref0, eps;
    D gratio[2];
    __device__ __forceinline__ void operator()(const D (&x)[3], D (&y)[2]) const {
        cloud_terminal_velocity<D>(pdf_c, vel, pref0, gratio, eps, x[0], x[1], x[2], y[0], y[1]);
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

template <class FT> D method_eps() { return is_f32<FT>() ? 1.1920928955078125e-07 : 2.220446049250313e-16; }
const D kPi = 3.141592653589793238462643383279502884;

template <class FT, class PDF, class VEL> RainVelSB make_rain_vel_sb(const PDF* pdf, const VEL* vel) {
    RainVelSB f{};
    if (!pdf || !vel) return f;
    widen(*pdf, f.pdf_r);
    widen(*vel, f.vel);
    f.pi_rho_w = kPi * f.pdf_r.rho_w;
    f.eps = method_eps<FT>();
    return f;
}
template <class FT, class PDF, class VEL> RainVelChen make_rain_vel_chen(const PDF* pdf, const VEL* vel) {
    RainVelChen f{};
    if (!pdf || !vel) return f;
    widen(*pdf, f.pdf_r);
    widen(*vel, f.vel);
    f.pi_rho_w = kPi * f.pdf_r.rho_w;
    f.eps = method_eps<FT>();
    return f;
}
template <class FT, class PDF, class VEL> CloudVel make_cloud_vel(const PDF* pdf, const VEL* vel) {
    CloudVel f{};
    if (!pdf || !vel) return f;
    widen(*pdf, f.pdf_c);
    widen(*vel, f.vel);
    const D t = 6.0 / f.vel.rho_w / kPi;
    f.pref0 = (1.0 / 18) * std::cbrt(t * t) * f.vel.grav / f.vel.nu_air;
    const D nu = f.pdf_c.nu_c, mu = f.pdf_c.mu_c;
    const D z = (nu + 1) / mu;
    f.gratio[0] = std::tgamma((nu + 1 + D(2.0 / 3)) / mu) / std::tgamma(z);
    f.gratio[1] = std::tgamma((nu + 1 + D(5.0 / 3)) / mu) / std::tgamma(z);
    f.eps = method_eps<FT>();
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
        return termvel_impl<FT>(pdf_r, vel, make_rain_vel_sb<FT>(pdf_r, vel), n, q_rai, rho, N_rai, vt0, vt1, stream,  \
                                "termvel_2m_rain_sb launch");                                                          \
    }                                                                                                                  \
    int cumicro_termvel_2m_rain_chen_##SUF(const cumicro_sb_pdf_r_##SUF* pdf_r,                                         \
                                           const cumicro_vel_chen_rain_##SUF* vel, int64_t n, const FT* q_rai,         \
                                           const FT* rho, const FT* N_rai, FT* vt0, FT* vt1, void* stream) {           \
        return termvel_impl<FT>(pdf_r, vel, make_rain_vel_chen<FT>(pdf_r, vel), n, q_rai, rho, N_rai, vt0, vt1,       \
                                stream, "termvel_2m_rain_chen launch");                                                \
    }                                                                                                                  \
    int cumicro_termvel_2m_cloud_##SUF(const cumicro_sb_pdf_c_##SUF* pdf_c, const cumicro_vel_stokes_##SUF* vel,        \
                                       int64_t n, const FT* q_lcl, const FT* rho, const FT* N_lcl, FT* vt0, FT* vt1,   \
                                       void* stream) {                                                                 \
        return termvel_impl<FT>(pdf_c, vel, make_cloud_vel<FT>(pdf_c, vel), n, q_lcl, rho, N_lcl, vt0, vt1, stream,   \
                                "termvel_2m_cloud launch");                                                            \
    }

CUMICRO_DEF_2M(f64, double)
CUMICRO_DEF_2M(f32, float)

int cumicro_rain_evaporation_2m_f64(const cumicro_params_2m_warm_f64* p, int64_t n, const double* const* in8, double* const* out4, void* stream) {
    return rain_evaporation_impl<double>(p, n, in8, out4, stream);
}
int cumicro_rain_evaporation_2m_f32(const cumicro_params_2m_warm_f32* p, int64_t n, const float* const* in8, float* const* out4, void* stream) {
    return rain_evaporation_impl<float>(p, n, in8, out4, stream);
}

}  // extern "C"
