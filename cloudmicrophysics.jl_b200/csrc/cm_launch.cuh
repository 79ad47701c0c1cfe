// cm_launch.cuh — the one streaming kernel shape every cumicro family uses.
//
// All hot-path functions are pointwise over structure-of-arrays columns: NIN input
// columns, NOUT output columns, n points, no coupling between points.  A family
// supplies a functor F (parameters by value, constant-bank resident through the
// __grid_constant__ kernel argument) with
//     __device__ void operator()(const double (&x)[NIN], double (&y)[NOUT]) const;
// The column type FT (double or float) is the I/O type only: arithmetic is always
// Float64 (a Float32 method loads float, computes in double with the Float32 method's
// regime thresholds, and rounds once on store — DESIGN.md §Float32);
// and this header supplies the data movement:
//   * vector variant: one thread owns VEC = 16 B / sizeof(FT) consecutive points, every
//     input column is read exactly once with one 128-bit ld.global.nc and every live output
//     column written exactly once with one 128-bit st.global.cs (streaming, never re-read);
//     scalar variant: one point per thread, 64/32-bit accesses, a warp covers one contiguous
//     256/128-byte run per column (used when registers, not bytes, are the scarce resource);
//   * all loads of an item are issued before any arithmetic, so 7-12 independent
//     128-bit requests per thread are in flight while the FP64 pipe works on the
//     previous item of other warps;
//   * the VEC points of one thread are evaluated in one unrolled body so the compiler
//     interleaves their (independent) FP64 dependency chains;
//   * grid-stride launch: blocks = SMs x resident blocks x 16, a whole number of waves on the
//     148 SMs (dynamic block scheduling evens out the SM-to-SM spread);
//   * a scalar variant covers misaligned columns and the n % VEC tail.
// NULL output pointers are skipped (optional diagnostics columns).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "cm_types.cuh"

namespace cm {

// a functor that calls log_abs_ declares `static constexpr bool kNeedsLog2 = true;` (its 4 KB table is staged only then)
template <class F, class = void> struct needs_log2 : std::false_type {};
template <class F> struct needs_log2<F, std::void_t<decltype(F::kNeedsLog2)>> : std::true_type {};
template <int BLOCK, class F> CM_DEV void math_tables_init_for() {
    math_tables_init<BLOCK>();
    if constexpr (needs_log2<F>::value) math_tables_init_log2<BLOCK>();
}

template <class FT, int NIN, int NOUT, class F> struct PointwiseArgs {
    F f;
    const FT* in[NIN];
    FT* out[NOUT];
    int64_t n;
};

template <class FT, int NIN, int NOUT, class F, bool VECTOR, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
pointwise_kernel(const __grid_constant__ PointwiseArgs<FT, NIN, NOUT, F> a, int64_t first) {
    constexpr int VEC = VECTOR ? vec<FT>::N : 1;
    math_tables_init_for<BLOCK, F>();  // exp/log tables -> shared memory (cm_math.cuh)
    const int64_t stride = (int64_t)gridDim.x * BLOCK;
    const int64_t n_items = VECTOR ? (a.n / VEC) : (a.n - first);
    for (int64_t it = (int64_t)blockIdx.x * BLOCK + threadIdx.x; it < n_items; it += stride) {
        const int64_t i0 = VECTOR ? it * VEC : first + it;
        double x[VEC][NIN];
        if constexpr (VECTOR) {
#pragma unroll
            for (int c = 0; c < NIN; ++c) {
                auto pk = ldg_vec(a.in[c] + i0);
#pragma unroll
                for (int v = 0; v < VEC; ++v) x[v][c] = (double)pk.v[v];
            }
        } else {
#pragma unroll
            for (int c = 0; c < NIN; ++c) x[0][c] = (double)__ldg(a.in[c] + i0);
        }
        double y[VEC][NOUT];
#pragma unroll
        for (int v = 0; v < VEC; ++v) a.f(x[v], y[v]);
        if constexpr (VECTOR) {
#pragma unroll
            for (int c = 0; c < NOUT; ++c) {
                if (a.out[c] == nullptr) continue;
                pack<FT, VEC> pk;
#pragma unroll
                for (int v = 0; v < VEC; ++v) pk.v[v] = (FT)y[v][c];
                st_vec(a.out[c] + i0, pk);
            }
        } else {
#pragma unroll
            for (int c = 0; c < NOUT; ++c)
                if (a.out[c]) a.out[c][i0] = (FT)y[0][c];
        }
    }
}


// Software-pipelined one-point-per-thread variant: the inputs of a thread's NEXT grid-stride item are fetched with
// per-thread asynchronous copies (cp.async / LDGSTS, global -> shared, no registers held) while the current item
// is computed, so the ~900-instruction FP64 body never waits on HBM latency (ncu: long-scoreboard stalls were the
// largest stall class of the register-loading variant).  Each thread reads back only its own slots: no block barrier.
template <class T> __device__ __forceinline__ void cp_async_elem(T* smem_dst, const T* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gmem_src), "n"(sizeof(T)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ALL_OUT: every output column is wanted (decided at launch): no per-column NULL test in the store sequence
constexpr size_t kPipelinedStaticLimit = 40 * 1024;
template <class FT, int NIN, int BLOCK> constexpr size_t pipelined_stage_bytes() { return sizeof(FT) * 2 * NIN * BLOCK; }
template <class FT, int NIN, int NOUT, class F, int BLOCK, int MINB, bool ALL_OUT = false>
__global__ void __launch_bounds__(BLOCK, MINB)
pointwise_kernel_pipelined(const __grid_constant__ PointwiseArgs<FT, NIN, NOUT, F> a) {
    math_tables_init_for<BLOCK, F>();
    // the staging area: static up to 40 KB, dynamic above (one large block per SM, e.g. 896 x 7 doubles x 2 = 100 KB)
    constexpr bool DYN = pipelined_stage_bytes<FT, NIN, BLOCK>() > kPipelinedStaticLimit;
    extern __shared__ __align__(16) unsigned char pipelined_dyn_smem[];
    __shared__ FT stage_static[DYN ? 1 : 2][DYN ? 1 : NIN][DYN ? 1 : BLOCK];
    FT (*stage)[NIN][BLOCK] = DYN ? reinterpret_cast<FT (*)[NIN][BLOCK]>(pipelined_dyn_smem) : reinterpret_cast<FT (*)[NIN][BLOCK]>(&stage_static[0][0][0]);
    const int tid = threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * BLOCK;
    int64_t it = (int64_t)blockIdx.x * BLOCK + tid;
    if (it < a.n) {
#pragma unroll
        for (int c = 0; c < NIN; ++c) cp_async_elem(&stage[0][c][tid], a.in[c] + it);
    }
    cp_async_commit();
    int buf = 0;
    for (; it < a.n; it += stride) {
        const int64_t nxt = it + stride;
        if (nxt < a.n) {
#pragma unroll
            for (int c = 0; c < NIN; ++c) cp_async_elem(&stage[buf ^ 1][c][tid], a.in[c] + nxt);
        }
        cp_async_commit();
        cp_async_wait<1>();   // everything but the group just committed has landed
        double x[NIN];
#pragma unroll
        for (int c = 0; c < NIN; ++c) x[c] = (double)stage[buf][c][tid];
        double y[NOUT];
        a.f(x, y);
#pragma unroll
        for (int c = 0; c < NOUT; ++c)
            if (ALL_OUT || a.out[c]) __stcs(a.out[c] + it, (FT)y[c]);
        buf ^= 1;
    }
}

// ---- tile variant: the launch shape of the headline kernel (cm_tile2m.cuh) for any functor ---------------------------------
// Block-uniform tile loop (tile = BLOCK consecutive points, dealt round-robin to the persistent blocks).  The NIN input tiles are
// fetched by 1-D bulk asynchronous copies (cp.async.bulk global -> shared, completion on an mbarrier, two tiles in flight) issued
// by one elected lane of up to NIN warps: no per-thread address arithmetic or load instruction, and the block barrier + the
// asynchronous copies inside the loop keep ptxas from hoisting the functor's constants into (spilled) uniform registers.
// Columns must be 16-byte aligned for the bulk copies; otherwise, and for a partial last tile, the same kernel uses guarded
// scalar loads (identical results).
namespace tile {
CM_DEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
CM_DEV void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
CM_DEV void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
CM_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
CM_DEV void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// one lane of a converged warp (the canonical leader election: ptxas keeps the guarded code on the uniform datapath)
CM_DEV bool elect_one() {
    unsigned p;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(p));
    return p != 0;
}
}  // namespace tile

template <class FT, int NIN, int NOUT, class F, int BLOCK, int MINB, bool ALL_OUT>
__global__ void __launch_bounds__(BLOCK, MINB) pointwise_kernel_tiled(const __grid_constant__ PointwiseArgs<FT, NIN, NOUT, F> a, int bulk_ok) {
    constexpr int NWARP = BLOCK / 32;
    extern __shared__ __align__(128) unsigned char tiled_dyn_smem[];
    FT (*stage)[NIN][BLOCK] = reinterpret_cast<FT (*)[NIN][BLOCK]>(tiled_dyn_smem);   // [2][NIN][BLOCK]
    __shared__ __align__(8) unsigned long long full[2];
    math_tables_init_for<BLOCK, F>();
    const int tid = threadIdx.x;
    const unsigned n_tiles = (unsigned)((a.n + BLOCK - 1) / BLOCK);
    const unsigned n_full = bulk_ok ? (unsigned)(a.n / BLOCK) : 0u;
    constexpr unsigned kTileBytes = BLOCK * sizeof(FT);
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction
    const bool issuer = warp < NIN;
    constexpr int kIssuers = NWARP < NIN ? NWARP : NIN;
    if (tid == 0) {
        tile::mbar_init(&full[0], kIssuers);
        tile::mbar_init(&full[1], kIssuers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](unsigned t, int buf) {
        if (t < n_full && tile::elect_one()) {
            const int ncol = (NIN - warp + NWARP - 1) / NWARP;
            tile::mbar_expect_tx(&full[buf], ncol * kTileBytes);
#pragma unroll 1
            for (int c = warp; c < NIN; c += NWARP) tile::bulk_g2s(&stage[buf][c][0], a.in[c] + (size_t)t * BLOCK, kTileBytes, &full[buf]);
        }
    };
    unsigned t = blockIdx.x;
    if (issuer) {
        issue(t, 0);
        if (t + gridDim.x < n_tiles) issue(t + gridDim.x, 1);
    }
    unsigned j = 0;
    for (; t < n_tiles; t += gridDim.x, ++j) {
        const int buf = j & 1;
        const int64_t it = (int64_t)t * BLOCK + tid;
        double x[NIN];
        if (t < n_full) {
            tile::mbar_wait(&full[buf], (j >> 1) & 1);
#pragma unroll
            for (int c = 0; c < NIN; ++c) x[c] = (double)stage[buf][c][tid];
        } else {
#pragma unroll
            for (int c = 0; c < NIN; ++c) x[c] = (double)__ldg(a.in[c] + (it < a.n ? it : a.n - 1));
        }
        __syncthreads();   // every thread has read stage[buf]: it may be refilled
        if (issuer) {
            const unsigned t2 = t + 2 * gridDim.x;
            if (t2 < n_tiles) issue(t2, buf);
        }
        if (it < a.n) {   // (a functor may count domain errors: a padding thread must not run it)
            double y[NOUT];
            a.f(x, y);
#pragma unroll
            for (int c = 0; c < NOUT; ++c)
                if (ALL_OUT || a.out[c]) __stcs(a.out[c] + it, (FT)y[c]);
        }
    }
}

template <class FT, int NIN, int NOUT, class F, int BLOCK, int MINB>
int launch_pointwise_tiled(const F& f, int64_t n, const FT* const (&in)[NIN], FT* const (&out)[NOUT], cudaStream_t stream, const char* what,
                           int waves = 8) {
    if (n == 0) return CUMICRO_OK;
    PointwiseArgs<FT, NIN, NOUT, F> a;
    a.f = f;
    a.n = n;
    int bulk_ok = 1;
    bool all_out = true;
    for (int c = 0; c < NIN; ++c) { a.in[c] = in[c]; if (!cmh::aligned16(in[c])) bulk_ok = 0; }
    for (int c = 0; c < NOUT; ++c) { a.out[c] = out[c]; all_out = all_out && out[c] != nullptr; }
    auto kern = all_out ? pointwise_kernel_tiled<FT, NIN, NOUT, F, BLOCK, MINB, true> : pointwise_kernel_tiled<FT, NIN, NOUT, F, BLOCK, MINB, false>;
    constexpr size_t smem = sizeof(FT) * 2 * NIN * BLOCK;
    // function attributes are per device: a process that drives several GPUs must not rely on a process-wide "already set" flag
    // for the one that decides whether the launch is legal (the carve-out is a hint and is set once)
    static bool hint_set[2] = {false, false};
    if (!hint_set[all_out]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        hint_set[all_out] = true;
    }
    if (smem > 48 * 1024 - 8 * 1024) {   // dynamic + the static math tables (up to 8 KB) against the 48 KB default
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cmh::cuda_status(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    }
#ifdef CUMICRO_TUNING
    { const char* wv0 = getenv("CUMICRO_WAVES"); if (wv0) waves = atoi(wv0); }
#endif
    const int64_t n_tiles = (n + BLOCK - 1) / BLOCK;
    const int blocks = (int)std::min<int64_t>(n_tiles, (int64_t)cmh::num_sms() * MINB * waves);
    kern<<<blocks, BLOCK, smem, stream>>>(a, bulk_ok);
    cmh::count_launch();
    return cmh::cuda_status(cudaGetLastError(), what);
}

// Enqueue F over n points on `stream`.  BLOCK x MINB fixes the register budget
// (65536 / (BLOCK*MINB) per thread).  USE_VEC = false launches the one-point-per-thread
// variant even for aligned columns: for the FP64-pipe-bound families (2M, 1M) the
// measured optimum is 32 resident warps/SM at 64 registers with 64-bit loads (a warp
// still reads one contiguous 256-byte run per column), tools/tune_2m.py.
template <class FT, int NIN, int NOUT, class F, int BLOCK = 256, int MINB = 1, bool USE_VEC = true, bool PIPELINED = false>
int launch_pointwise(const F& f, int64_t n, const FT* const (&in)[NIN], FT* const (&out)[NOUT], cudaStream_t stream,
                     const char* what) {
    if (n == 0) return CUMICRO_OK;
    PointwiseArgs<FT, NIN, NOUT, F> a;
    a.f = f;
    a.n = n;
    bool vec_ok = USE_VEC;
    for (int c = 0; c < NIN; ++c) { a.in[c] = in[c]; vec_ok = vec_ok && cmh::aligned16(in[c]); }
    for (int c = 0; c < NOUT; ++c) { a.out[c] = out[c]; vec_ok = vec_ok && (out[c] == nullptr || cmh::aligned16(out[c])); }
    constexpr int VEC = vec<FT>::N;
#ifdef CUMICRO_TUNING
    { static const char* fs = getenv("CUMICRO_FORCE_SCALAR"); if (fs && fs[0] == '1') vec_ok = false; }
#endif
    // 16 waves of the resident grid: blocks are dealt to SMs as they finish, which evens out the SM-to-SM spread
    // (measured on the 2M kernel: 2 waves 0.699 ms, 4: 0.666, 16: 0.643, 32: 0.633, one item per thread 0.659)
    int waves = 16;
#ifdef CUMICRO_TUNING
    { static const char* wv0 = getenv("CUMICRO_WAVES"); if (wv0) waves = atoi(wv0); }
#endif
    const int max_blocks = cmh::num_sms() * MINB * waves;
    if constexpr (PIPELINED && !USE_VEC) {
        bool ok = true;   // cp.async needs naturally aligned elements (always true for FT columns) -- keep the plain path for odd pointers
        for (int c = 0; c < NIN; ++c) ok = ok && (reinterpret_cast<uintptr_t>(in[c]) % sizeof(FT) == 0);
#ifdef CUMICRO_TUNING
        { static const char* np = getenv("CUMICRO_NO_PIPELINE"); if (np && np[0] == '1') ok = false; }
#endif
        if (ok) {
            bool all_out = true;
            for (int c = 0; c < NOUT; ++c) all_out = all_out && out[c] != nullptr;
            auto kern = all_out ? pointwise_kernel_pipelined<FT, NIN, NOUT, F, BLOCK, MINB, true>
                                : pointwise_kernel_pipelined<FT, NIN, NOUT, F, BLOCK, MINB, false>;
            static bool carveout_set[2] = {false, false};   // 2 x NIN x BLOCK staged elements + the math tables per block, MINB blocks per SM
            if (!carveout_set[all_out]) {
                cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                carveout_set[all_out] = true;
            }
            const int blocks = (int)std::min<int64_t>((n + BLOCK - 1) / BLOCK, (int64_t)max_blocks);
            constexpr size_t stage_bytes = pipelined_stage_bytes<FT, NIN, BLOCK>();
            constexpr size_t dyn = stage_bytes > kPipelinedStaticLimit ? stage_bytes : 0;
            if (dyn) {   // every launch: function attributes are per device
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
                if (e != cudaSuccess) return cmh::cuda_status(e, "pipelined: cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
            }
            kern<<<blocks, BLOCK, dyn, stream>>>(a);
            cmh::count_launch();
            return cmh::cuda_status(cudaGetLastError(), what);
        }
    }
    int64_t first = 0;
    if (vec_ok && n >= VEC) {
        int64_t items = n / VEC;
        int blocks = (int)std::min<int64_t>((items + BLOCK - 1) / BLOCK, max_blocks);
        pointwise_kernel<FT, NIN, NOUT, F, true, BLOCK, MINB><<<blocks, BLOCK, 0, stream>>>(a, 0);
        cmh::count_launch();
        first = items * VEC;
    }
    if (first < n) {
        int64_t items = n - first;
        int blocks = (int)std::min<int64_t>((items + BLOCK - 1) / BLOCK, max_blocks);
        pointwise_kernel<FT, NIN, NOUT, F, false, BLOCK, MINB><<<blocks, BLOCK, 0, stream>>>(a, first);
        cmh::count_launch();
    }
    return cmh::cuda_status(cudaGetLastError(), what);
}

// Shared argument validation of the C-ABI entry points.
template <class FT, int NIN>
int validate_columns(const void* params, int64_t n, const FT* const (&in)[NIN]) {
    if (params == nullptr) return cmh::fail(CUMICRO_E_NULL, "parameter block is NULL");
    if (n < 0) return cmh::fail(CUMICRO_E_SIZE, "n = %lld is negative", (long long)n);
    for (int c = 0; c < NIN; ++c)
        if (n > 0 && in[c] == nullptr) return cmh::fail(CUMICRO_E_NULL, "input column %d is NULL", c);
    return CUMICRO_OK;
}
template <class FT, int NOUT> int require_outputs(int64_t n, FT* const (&out)[NOUT], int n_required) {
    for (int c = 0; c < n_required; ++c)
        if (n > 0 && out[c] == nullptr) return cmh::fail(CUMICRO_E_NULL, "output column %d is NULL", c);
    return CUMICRO_OK;
}

}  // namespace cm
