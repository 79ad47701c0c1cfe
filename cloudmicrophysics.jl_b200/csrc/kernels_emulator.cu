// kernels_emulator.cu — the trained-emulator methods of AerosolActivation (ext/EmulatorModelsExt.jl:32-103) for a
// multilayer-perceptron machine: cumicro_aa_emulated_* and cumicro_emulator_weight_count_* (include/cumicro.h).
//
// A ROW is one (grid point, mode i) pair: the reference builds the feature row with modes 1 and i swapped and asks the machine
// for the activated fraction of "mode 1".  A block owns the rows of floor(32 / n_modes) consecutive points.  The rows'
// activations live in shared memory as
// [unit][row] (a warp reads one unit of 32 rows conflict-free; all threads of a layer read the same [unit][*] line, a broadcast),
// a thread owns one output unit of a layer and keeps its row accumulators in registers — all 32 rows for a layer at least 129
// units wide, 32 / G rows when G = 2, 4, 8, 16 groups of threads fit a narrower layer into the block — and the weights stream
// through L1/L2 as [in][out], out fastest: consecutive threads read consecutive words.  Per input unit a thread issues 1 global
// load, R / 2 shared 128-bit loads and R DFMAs; the sums run in input order (the order of a plain
// dot product), in Float64 for both float types.
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
    int max_width;    // widest layer input or output: the activation buffers are [max_width][kRows] each
};

__device__ __forceinline__ double emu_act(int kind, double x) {
    switch (kind) {
        case 0: return (x > 0.0) ? x : ((x != x) ? x : 0.0);   // relu, NaN kept like max(0, NaN) in the reference's language
        case 1: return tanh(x);
        case 2: return 1.0 / (1.0 + exp(-x));
        default: return x;
    }
}

// One dense layer of a tile: thread t owns output unit h = t mod H' of the R = 32 / G rows [g R, g R + R), g = t / H' (H' = H for
// G > 1; for G = 1 a thread walks h = t, t + 256, ...).  W as [K][H], H fastest; activations [unit][32 rows].
template <class FT, int R>
__device__ __forceinline__ void dense(const FT* __restrict__ W, const FT* __restrict__ B, const double* __restrict__ in, double* __restrict__ out,
                                      int K, int H, bool last, int activation) {
    constexpr int G = kRows / R;
    const int g = (G == 1) ? 0 : threadIdx.x / H;
    const int h0 = (G == 1) ? threadIdx.x : threadIdx.x - g * H;
    if (g >= G) return;
    const int r0 = g * R;
    for (int h = h0; h < H; h += (G == 1 ? kThreads : H)) {
        double acc[R];
        const double b = (double)__ldg(B + h);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = b;
        // weights of U input units are fetched ahead of their multiply-adds (an L1/L2 hit is ~10x the 2 R issue cycles of a unit)
        constexpr int U = (R >= 32) ? 4 : 8;
        int k = 0;
        for (; k + U <= K; k += U) {
            double wv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) wv[u] = (double)__ldg(W + (size_t)(k + u) * H + h);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double2* x2 = reinterpret_cast<const double2*>(in + (k + u) * kRows + r0);
#pragma unroll
                for (int r = 0; r < R / 2; ++r) {
                    const double2 v = x2[r];
                    acc[2 * r] = fma(wv[u], v.x, acc[2 * r]);
                    acc[2 * r + 1] = fma(wv[u], v.y, acc[2 * r + 1]);
                }
            }
        }
        for (; k < K; ++k) {
            const double wv = (double)__ldg(W + (size_t)k * H + h);
            const double2* x2 = reinterpret_cast<const double2*>(in + k * kRows + r0);
#pragma unroll
            for (int r = 0; r < R / 2; ++r) {
                const double2 v = x2[r];
                acc[2 * r] = fma(wv, v.x, acc[2 * r]);
                acc[2 * r + 1] = fma(wv, v.y, acc[2 * r + 1]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) out[h * kRows + r0 + r] = last ? acc[r] : emu_act(activation, acc[r]);
        if (G != 1) break;
    }
}

template <class FT>
__global__ void __launch_bounds__(kThreads, 2) emu_kernel(const __grid_constant__ EmuArgs<FT> a) {
    extern __shared__ double smem[];
    double* buf0 = smem;
    double* buf1 = smem + (size_t)a.max_width * kRows;
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
            buf0[f * kRows + r] = x;
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
            // narrow layers: the 32 rows are split over G = 2^g thread groups (G H <= 256), each thread keeps 32 / G accumulators
            int G = 1;
            while (G < kRows && 2 * G * H <= kThreads) G *= 2;
            switch (G) {
                case 1: dense<FT, 32>(W, B, in, out, K, H, last, a.p.activation); break;
                case 2: dense<FT, 16>(W, B, in, out, K, H, last, a.p.activation); break;
                case 4: dense<FT, 8>(W, B, in, out, K, H, last, a.p.activation); break;
                case 8: dense<FT, 4>(W, B, in, out, K, H, last, a.p.activation); break;
                default: dense<FT, 2>(W, B, in, out, K, H, last, a.p.activation); break;   // G = 16 (and 32: half the groups idle)
            }
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
    int mw = 4 * p->n_modes + 3;
    for (int l = 0; l < p->n_layers; ++l) mw = std::max(mw, (int)p->width[l]);
    a.max_width = mw;
    const size_t shmem = sizeof(double) * 2 * (size_t)mw * kRows;
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
