// kernels_probe.cu — roofline denominators measured in place.
//
// The tendency kernels are bound by the FP64 pipe, whose peak is not in
// MEASURED_PEAKS.json (that file holds the HBM copy rate and the bf16 tensor rate).
// cumicro_probe_fp64_fma runs a register-resident chain of independent DFMAs so
// bench.py can time the FP64 FMA peak of the very GPU (and clocks) the bench runs on.
#include "cm_launch.cuh"

namespace {

constexpr int kChains = 8;

__global__ void __launch_bounds__(256) fp64_fma_probe(double* out, int64_t iters, double a, double b) {
    double x[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) x[k] = 1.0 + 1e-9 * (threadIdx.x + k);
    for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += x[k];
    if (s == 12345.678) out[0] = s;  // never true; keeps the chain alive
}

// The device math of cm_math.cuh as a pointwise family, so tests can measure its
// accuracy on the GPU itself (hardware MUFU seeds) against mpmath.
struct MathProbe {
    int fn;
    __device__ __forceinline__ void operator()(const double (&x)[2], double (&y)[1]) const {
        switch (fn) {
            case 0: y[0] = cm::exp_(x[0]); break;
            case 1: y[0] = cm::logp_(x[0]); break;
            case 2: y[0] = cm::cbrtp_(x[0]); break;
            case 3: y[0] = cm::rcp_(x[0]); break;
            case 4: y[0] = cm::powp_(x[0], x[1]); break;
            case 5: y[0] = cm::exp_full_(x[0]); break;
            case 6: y[0] = cm::sqrtp_(x[0]); break;
            case 7: y[0] = cm::erf_(x[0]); break;
            case 8: y[0] = cm::log1p_pos_(x[0]); break;
            case 9: y[0] = cm::rcbrtp_(x[0]); break;
            case 10: y[0] = cm::divr_(x[0], x[1], cm::rcp_cr_(x[1])); break;
            case 11: y[0] = cm::erf_fast_(x[0]); break;
            default: y[0] = 0.0;
        }
    }
};

}  // namespace

extern "C" {

int cumicro_probe_math_f64(int fn, int64_t n, const double* x, const double* y, double* out, void* stream) {
    const double* in[2] = {x, y};
    double* o[1] = {out};
    int st = cm::validate_columns<double, 2>(&fn, n, in);
    if (st) return st;
    if ((st = cm::require_outputs<double, 1>(n, o, 1))) return st;
    if (fn < 0 || fn > 11) return cmh::fail(CUMICRO_E_ARG, "probe_math: unknown function id %d", fn);
    return cm::launch_pointwise<double, 2, 1, MathProbe, 256, 2>(MathProbe{fn}, n, in, o, (cudaStream_t)stream,
                                                                 "math probe launch");
}

int cumicro_probe_fp64_fma(int64_t iters, int blocks_per_sm, double* scratch, double* flops_out, void* stream) {
    if (scratch == nullptr || flops_out == nullptr) return cmh::fail(CUMICRO_E_NULL, "probe: NULL pointer");
    if (iters <= 0 || blocks_per_sm <= 0) return cmh::fail(CUMICRO_E_ARG, "probe: iters and blocks_per_sm must be > 0");
    const int blocks = cmh::num_sms() * blocks_per_sm;
    fp64_fma_probe<<<blocks, 256, 0, (cudaStream_t)stream>>>(scratch, iters, 0.999999999, 1e-9);
    cmh::count_launch();
    *flops_out = 2.0 * kChains * (double)iters * 256.0 * (double)blocks;
    return cmh::cuda_status(cudaGetLastError(), "fp64 probe launch");
}

}  // extern "C"
