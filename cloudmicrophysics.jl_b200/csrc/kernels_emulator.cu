// kernels_emulator.cu — the trained-emulator methods of AerosolActivation (ext/EmulatorModelsExt.jl:32-103) for a
// multilayer-perceptron machine: cumicro_aa_emulated_* and cumicro_emulator_weight_count_* (include/cumicro.h).
//
// A ROW is one (grid point, mode i) pair: the reference builds the feature row with modes 1 and i swapped and asks the machine
// for the activated fraction of "mode 1".  A block owns the rows of floor(32 / n_modes) consecutive points.  The rows'
// activations live in shared memory as [unit][row]; every dense layer is a small GEMM, Out[32][H] = In[32][K] W[K][H], and runs
// on the FP64 tensor cores (dense_mma below) — the one GEMM-shaped operation of this library.  Float64 accumulation for both
// float types; the sums run in the tensor core's order (rounding-level differences from a plain dot product).
#include <algorithm>

#include "cm_launch.cuh"

namespace {

using namespace cm;

constexpr int kRows = 32;        // rows per block tile
constexpr int kThreads = 256;    // = the widest layer
constexpr int kMaxWidth = 256;
constexpr int kMaxLayers = 4;
constexpr int kMaxFeat = 35;

template <class FT> struct EmuArgs {
    cumicro_params_emulator_f64 p;
    const FT* weights;
    const FT* T;
    const FT* p_air;
    const FT* w;
    FT* N_act[8];
    FT* N_tot;
    int64_t n;        // grid points
    int width_a, width_b;   // the two activation buffers, [width][kRS] each: A holds the features and the outputs of layers 1, 3 (0-based), B those of layers 0, 2
};

__device__ __forceinline__ double emu_act(int kind, double x) {
    switch (kind) {
        case 0: return (x > 0.0) ? x : ((x != x) ? x : 0.0);   // relu, NaN kept like max(0, NaN) in the reference's language
        case 1: return tanh(x);
        case 2: return 1.0 / (1.0 + exp(-x));
        default: return x;
    }
}

// One dense layer of a tile on the FP64 tensor cores: Out[32 rows][H] = In[32 rows][K] W[K][H] + b as 8x8x4 warp-level
// matrix multiply-accumulates (mma.sync.aligned.m8n8k4 f64: the only FP64 tensor-core shape; tcgen05 has no FP64 kind).
// Fragment coordinates of lane T: A[row = T/4][k = T%4], B[k = T%4][unit = T/4], C[row = T/4][unit = 2 (T%4) + {0, 1}].
// A warp owns two 8-unit groups per pass (ug, ug + 8) for all four 8-row groups: per step of four input units it loads
// 4 A fragments (shared memory, [unit][kRS] with kRS = 40: the four k-lanes of an 8-byte fragment load fall into disjoint bank
// halves) and 2 B fragments (the weights, 64-byte runs through L1/L2, fetched four steps ahead) for 8 MMAs = 2048
// multiply-adds.  Inputs / units beyond K / H enter as zeros.  The bias seeds the accumulators; hidden layers apply the activation on the way out.
constexpr int kRS = 40;
__device__ __forceinline__ void dmma8x8x4(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <class FT>
__device__ __forceinline__ void dense_mma(const FT* __restrict__ W, const FT* __restrict__ B, const double* __restrict__ in,
                                          double* __restrict__ out, int K, int H, bool last, int activation) {
    constexpr int NW = kThreads / 32, UG = 2, PF = 4;   // unit groups per warp and pass; steps of weights fetched ahead
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lr = lane >> 2, lk = lane & 3;
    const int n_ug = (H + 7) >> 3;
    for (int ug0 = warp; ug0 < n_ug; ug0 += NW * UG) {      // this warp's groups of the pass: ug0, ug0 + NW
        double acc[UG][4][2];
#pragma unroll
        for (int q = 0; q < UG; ++q) {
            const int u0 = (ug0 + NW * q) * 8 + 2 * lk;
            const double b0 = (u0 < H) ? (double)__ldg(B + u0) : 0.0, b1 = (u0 + 1 < H) ? (double)__ldg(B + u0 + 1) : 0.0;
#pragma unroll
            for (int rg = 0; rg < 4; ++rg) { acc[q][rg][0] = b0; acc[q][rg][1] = b1; }
        }
        const bool two = ug0 + NW < n_ug;                    // warp-uniform: the second group exists
        for (int k0 = 0; k0 < K; k0 += 4 * PF) {
            // the weights of PF steps first (L1 / L2 latency is ~10 steps of MMAs), then step by step: A fragments + MMAs
            double bfr[PF][UG];
#pragma unroll
            for (int st = 0; st < PF; ++st) {
                const int kk = k0 + 4 * st + lk;
#pragma unroll
                for (int q = 0; q < UG; ++q) {
                    const int unit = (ug0 + NW * q) * 8 + lr;
                    bfr[st][q] = (kk < K && unit < H) ? (double)__ldg(W + (size_t)kk * H + unit) : 0.0;
                }
            }
#pragma unroll
            for (int st = 0; st < PF; ++st) {
                const int kk = k0 + 4 * st + lk;
                if (k0 + 4 * st < K) {                       // warp-uniform
                    double afr[4];
#pragma unroll
                    for (int rg = 0; rg < 4; ++rg) afr[rg] = (kk < K) ? in[kk * kRS + rg * 8 + lr] : 0.0;
#pragma unroll
                    for (int rg = 0; rg < 4; ++rg) dmma8x8x4(acc[0][rg], afr[rg], bfr[st][0]);
                    if (two) {
#pragma unroll
                        for (int rg = 0; rg < 4; ++rg) dmma8x8x4(acc[1][rg], afr[rg], bfr[st][1]);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < UG; ++q) {
            const int ug = ug0 + NW * q;
            if (ug < n_ug) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int u = ug * 8 + 2 * lk + e;
                    if (u < H) {
#pragma unroll
                        for (int rg = 0; rg < 4; ++rg) {
                            const double v = acc[q][rg][e];
                            out[u * kRS + rg * 8 + lr] = last ? v : emu_act(activation, v);
                        }
                    }
                }
            }
        }
    }
}

template <class FT>
__global__ void __launch_bounds__(kThreads, 2) emu_kernel(const __grid_constant__ EmuArgs<FT> a) {
    extern __shared__ double smem[];
    double* buf0 = smem;
    double* buf1 = smem + (size_t)a.width_a * kRS;
    const int nm = a.p.n_modes;
    const int nfeat = 4 * nm + 3;
    const int ppt = kRows / nm;                       // whole points per tile: rows [0, ppt nm) are used, the rest idle
    const int rows_used = ppt * nm;
    const int64_t n_tiles = (a.n + ppt - 1) / ppt;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t pt0 = tile * ppt;
        // ---- the feature rows (ext/EmulatorModelsExt.jl:47-66), preprocessed (ext/Common.jl:57-77) and standardized
        for (int t = threadIdx.x; t < nfeat * kRows; t += kThreads) {
            const int f = t / kRows, r = t - f * kRows;
            const int lp = r / nm, i = r - lp * nm;
            const int64_t pt = pt0 + lp;
            double x = 0.0;
            if (r < rows_used && pt < a.n) {
                bool logged;
                if (f < 4 * nm) {
                    const int j = f >> 2, c = f & 3;
                    const int m = (j == 0) ? i : ((j == i) ? 0 : j);          // modes_perm[[1, i]] = modes_perm[[i, 1]]
                    x = (c == 0) ? a.p.mode_N[m] : ((c == 1) ? a.p.mode_mean[m] : ((c == 2) ? a.p.mode_stdev[m] : a.p.mode_kappa[m]));
                    logged = c < 2;
                } else {
                    const int c = f - 4 * nm;
                    x = (c == 0) ? (double)__ldg(a.w + pt) : ((c == 1) ? (double)__ldg(a.T + pt) : (double)__ldg(a.p_air + pt));
                    logged = c == 0;
                }
                if (a.p.log_features && logged) x = log(x);
                x = (x - a.p.feat_mean[f]) * a.p.feat_inv_scale[f];
            }
            buf0[f * kRS + r] = x;
        }
        __syncthreads();
        // ---- dense layers
        double* in = buf0;
        double* out = buf1;
        const FT* W = a.weights;
        int K = nfeat;
        for (int l = 0; l < a.p.n_layers; ++l) {
            const int H = a.p.width[l];
            const FT* B = W + (size_t)K * H;
            const bool last = l == a.p.n_layers - 1;
            dense_mma<FT>(W, B, in, out, K, H, last, a.p.activation);
            __syncthreads();
            W = B + H;
            K = H;
            double* t = in; in = out; out = t;
        }
        // ---- fraction -> activated number (ext/EmulatorModelsExt.jl:67); the optional sum in mode order (:93-103)
        {
            const int r = threadIdx.x;
            const int lp = r / nm, i = r - lp * nm;
            const int64_t pt = pt0 + lp;
            const bool live = r < rows_used && pt < a.n;
            if (live) {
                double y = in[r];
                if (a.p.target_transform) y = (1.0 / (2.0 * 0.99)) * tanh(y) + 0.5;
                const double frac = (y != y) ? y : fmax(0.0, fmin(1.0, y));
                const double N = frac * a.p.mode_N[i];
                out[r] = N;      // (the other buffer is free now) for the per-point sum below
                if (a.N_act[i]) a.N_act[i][pt] = (FT)N;
            }
            if (a.N_tot) {
                __syncthreads();
                if (live && i == 0) {
                    double s = out[r];
                    for (int j = 1; j < nm; ++j) s += out[r + j];
                    a.N_tot[pt] = (FT)s;
                }
            }
        }
        __syncthreads();
    }
}

template <class FT> struct PEmu;
template <> struct PEmu<double> { using type = cumicro_params_emulator_f64; };
template <> struct PEmu<float> { using type = cumicro_params_emulator_f32; };

template <class P> int emu_check(const P* p) {
    if (p == nullptr) return cmh::fail(CUMICRO_E_NULL, "parameter block is NULL");
    if (p->n_modes < 1 || p->n_modes > 8) return cmh::fail(CUMICRO_E_OPTION, "emulator: n_modes = %d (expected 1..8)", (int)p->n_modes);
    if (p->n_layers < 1 || p->n_layers > kMaxLayers) return cmh::fail(CUMICRO_E_OPTION, "emulator: n_layers = %d (expected 1..%d)", (int)p->n_layers, kMaxLayers);
    for (int l = 0; l < p->n_layers; ++l)
        if (p->width[l] < 1 || p->width[l] > kMaxWidth) return cmh::fail(CUMICRO_E_OPTION, "emulator: width[%d] = %d (expected 1..%d)", l, (int)p->width[l], kMaxWidth);
    if (p->width[p->n_layers - 1] != 1) return cmh::fail(CUMICRO_E_OPTION, "emulator: the last layer must have width 1, not %d", (int)p->width[p->n_layers - 1]);
    if (p->activation < 0 || p->activation > 3) return cmh::fail(CUMICRO_E_OPTION, "emulator: activation = %d (0 relu, 1 tanh, 2 logistic, 3 identity)", (int)p->activation);
    if ((p->log_features | 1) != 1 || (p->target_transform | 1) != 1) return cmh::fail(CUMICRO_E_OPTION, "emulator: log_features / target_transform must be 0 or 1");
    return CUMICRO_OK;
}

template <class P> int64_t emu_count(const P* p) {
    if (emu_check(p) != CUMICRO_OK) return -1;
    int64_t c = 0;
    int K = 4 * p->n_modes + 3;
    for (int l = 0; l < p->n_layers; ++l) { c += (int64_t)K * p->width[l] + p->width[l]; K = p->width[l]; }
    return c;
}

template <class FT>
int emu_launch(const typename PEmu<FT>::type* p, const FT* weights, int64_t n, const FT* T, const FT* p_air, const FT* w, FT* const* N_act,
               FT* N_tot, void* stream) {
    int st = emu_check(p);
    if (st) return st;
    if (n < 0) return cmh::fail(CUMICRO_E_SIZE, "n = %lld is negative", (long long)n);
    if (n == 0) return CUMICRO_OK;
    if (!weights) return cmh::fail(CUMICRO_E_NULL, "emulator: the weight buffer is NULL");
    if (!T || !p_air || !w) return cmh::fail(CUMICRO_E_NULL, "emulator: an input column (T, p, w) is NULL");
    if (!N_act && !N_tot) return cmh::fail(CUMICRO_E_NULL, "emulator: no output requested");
    EmuArgs<FT> a{};
    widen(*p, a.p);
    a.weights = weights; a.T = T; a.p_air = p_air; a.w = w; a.N_tot = N_tot; a.n = n;
    for (int i = 0; i < 8; ++i) a.N_act[i] = (N_act && i < p->n_modes) ? N_act[i] : nullptr;
    int wa = 4 * p->n_modes + 3, wb = 1;
    for (int l = 0; l < p->n_layers; ++l) {
        if (l & 1) wa = std::max(wa, (int)p->width[l]);
        else wb = std::max(wb, (int)p->width[l]);
    }
    a.width_a = wa;
    a.width_b = wb;
    const size_t shmem = sizeof(double) * (size_t)(wa + wb) * kRS;
    auto kern = emu_kernel<FT>;
    if (shmem > 48 * 1024) {
        st = cmh::cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem), "cudaFuncSetAttribute");
        if (st) return st;
    }
    const int ppt = kRows / p->n_modes;
    const int64_t tiles = (n + ppt - 1) / ppt;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(shmem, 1)));
    const int blocks = (int)std::min<int64_t>(tiles, (int64_t)cmh::num_sms() * per_sm);
    kern<<<blocks, kThreads, shmem, (cudaStream_t)stream>>>(a);
    cmh::count_launch();
    return cmh::cuda_status(cudaGetLastError(), "aa_emulated launch");
}

}  // namespace

extern "C" {

int64_t cumicro_emulator_weight_count_f64(const cumicro_params_emulator_f64* p) { return emu_count(p); }
int64_t cumicro_emulator_weight_count_f32(const cumicro_params_emulator_f32* p) { return emu_count(p); }

int cumicro_aa_emulated_f64(const cumicro_params_emulator_f64* p, const double* weights, int64_t n, const double* T, const double* p_air,
                            const double* w, double* const* N_act, double* N_tot, void* stream) {
    return emu_launch<double>(p, weights, n, T, p_air, w, N_act, N_tot, stream);
}
int cumicro_aa_emulated_f32(const cumicro_params_emulator_f32* p, const float* weights, int64_t n, const float* T, const float* p_air,
                            const float* w, float* const* N_act, float* N_tot, void* stream) {
    return emu_launch<float>(p, weights, n, T, p_air, w, N_act, N_tot, stream);
}

}  // extern "C"
