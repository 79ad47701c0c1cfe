// kernels_probe.cu — roofline denominators measured in place.
//
// The tendency kernels are bound by the FP64 pipe, whose peak is not in
// MEASURED_PEAKS.json (that file holds the HBM copy rate and the bf16 tensor rate).
// cumicro_probe_fp64_fma runs a register-resident chain of independent DFMAs so
// bench.py can time the FP64 FMA peak of the very GPU (and clocks) the bench runs on.
#include "cm_types.cuh"

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

}  // namespace

extern "C" {

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
