// cm_collective.cu — the only exchange step of the path: domain diagnostics of the column slabs.
//
//   cumicro_reduce_diagnostics_*  one slab's sums  out[k] = Σ_i w[i] cols[k][i]  (w = rho or NULL = 1), Float64 accumulation in a
//                                 fixed order (bit-reproducible, independent of the launch shape): the diagnostics of tendencies
//                                 that were computed by the non-fused entry points (the fused config-5 kernel reduces in-kernel).
//   cumicro_nccl_allreduce_f64    sum of `count` doubles over the ranks of the caller's ncclComm_t, in place, on the caller's stream.
//   cumicro_nccl_{unique_id,comm_init_rank,comm_destroy}   for hosts without an NCCL binding of their own.
// libcumicro.so has no link-time dependency on NCCL: the symbols are resolved at first use from the NCCL already loaded in the process
// (torch / NCCL.jl / MPI stack) or from libnccl.so.2 on the loader path.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "cm_hostpipe.cuh"

namespace {

constexpr int kRedBlock = 256;
constexpr int kRedBlocks = 592;   // 4 per SM on a B200; the partial order is fixed by (block, thread), not by timing
constexpr int kMaxCols = 16;

template <class FT> struct RedArgs {
    const FT* w;
    const FT* cols[kMaxCols];
    int ncols;
    int64_t n;
    double* partials;   // [kRedBlocks][ncols]
    double* out;
    unsigned int* ticket;
};

// Each block sums a fixed, contiguous-strided subset in a fixed order; the last block to finish (atomic ticket) adds the block
// partials in block order, so the result does not depend on scheduling.
template <class FT> __global__ void __launch_bounds__(kRedBlock) reduce_diag_kernel(const __grid_constant__ RedArgs<FT> a) {
    __shared__ double sh[kRedBlock / 32];
    __shared__ bool last;
    const int64_t stride = (int64_t)gridDim.x * kRedBlock;
    for (int k = 0; k < a.ncols; ++k) {
        double acc = 0.0;
        for (int64_t i = (int64_t)blockIdx.x * kRedBlock + threadIdx.x; i < a.n; i += stride) {
            const double v = (double)a.cols[k][i];
            acc += a.w ? (double)a.w[i] * v : v;
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < kRedBlock / 32; ++w) s += sh[w];
            a.partials[(size_t)blockIdx.x * a.ncols + k] = s;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        for (int k = threadIdx.x; k < a.ncols; k += kRedBlock) {
            double s = 0.0;
            for (unsigned b = 0; b < gridDim.x; ++b) s += a.partials[(size_t)b * a.ncols + k];
            a.out[k] = s;
        }
        if (threadIdx.x == 0) *a.ticket = 0u;   // ready for the next call on this stream
    }
}

template <class FT>
int reduce_impl(int64_t n, const FT* w, const FT* const* cols, int ncols, double* out, void* scratch, int64_t scratch_bytes, void* stream) {
    if (n < 0) return cmh::fail(CUMICRO_E_SIZE, "n = %lld is negative", (long long)n);
    if (ncols < 1 || ncols > kMaxCols) return cmh::fail(CUMICRO_E_ARG, "ncols = %d (expected 1..%d)", ncols, kMaxCols);
    if (cols == nullptr || out == nullptr || scratch == nullptr) return cmh::fail(CUMICRO_E_NULL, "cols / out / scratch is NULL");
    const int64_t need = (int64_t)sizeof(double) * kRedBlocks * ncols + 16;
    if (scratch_bytes < need) return cmh::fail(CUMICRO_E_SIZE, "scratch holds %lld bytes, %lld needed (cumicro_reduce_diagnostics_scratch_bytes)", (long long)scratch_bytes, (long long)need);
    RedArgs<FT> a{};
    a.w = w; a.ncols = ncols; a.n = n; a.out = out;
    for (int k = 0; k < ncols; ++k) {
        if (n > 0 && cols[k] == nullptr) return cmh::fail(CUMICRO_E_NULL, "column %d is NULL", k);
        a.cols[k] = cols[k];
    }
    // caller-owned scratch: [ticket (16 bytes, must be zero before the first use; the kernel leaves it zero)] [partials]
    a.ticket = static_cast<unsigned int*>(scratch);
    a.partials = reinterpret_cast<double*>(static_cast<char*>(scratch) + 16);
    reduce_diag_kernel<FT><<<kRedBlocks, kRedBlock, 0, (cudaStream_t)stream>>>(a);
    cmh::count_launch();
    return cmh::cuda_status(cudaGetLastError(), "reduce_diagnostics launch");
}

// ---- NCCL through dlsym -------------------------------------------------------------------------------------------------
struct NcclId { char internal[128]; };
using ncclAllReduce_t = int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t);
using ncclGetUniqueId_t = int (*)(NcclId*);
using ncclCommInitRank_t = int (*)(void**, int, NcclId, int);
using ncclCommDestroy_t = int (*)(void*);
using ncclGetErrorString_t = const char* (*)(int);
struct NcclApi {
    ncclAllReduce_t all_reduce = nullptr;
    ncclGetUniqueId_t unique_id = nullptr;
    ncclCommInitRank_t init_rank = nullptr;
    ncclCommDestroy_t destroy = nullptr;
    ncclGetErrorString_t err = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

const NcclApi& nccl() {
    std::call_once(g_nccl_once, [] {
        void* h = dlopen(nullptr, RTLD_NOW);                                  // already in the process (torch, NCCL.jl, ...)
        if (h == nullptr || dlsym(h, "ncclAllReduce") == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h == nullptr) return;
        g_nccl.all_reduce = (ncclAllReduce_t)dlsym(h, "ncclAllReduce");
        g_nccl.unique_id = (ncclGetUniqueId_t)dlsym(h, "ncclGetUniqueId");
        g_nccl.init_rank = (ncclCommInitRank_t)dlsym(h, "ncclCommInitRank");
        g_nccl.destroy = (ncclCommDestroy_t)dlsym(h, "ncclCommDestroy");
        g_nccl.err = (ncclGetErrorString_t)dlsym(h, "ncclGetErrorString");
        g_nccl.ok = g_nccl.all_reduce && g_nccl.unique_id && g_nccl.init_rank && g_nccl.destroy;
    });
    return g_nccl;
}
int nccl_status(int r, const char* what) {
    if (r == 0) return CUMICRO_OK;
    return cmh::fail(CUMICRO_E_ARG, "%s: NCCL error %d (%s)", what, r, g_nccl.err ? g_nccl.err(r) : "?");
}
int need_nccl() {
    return nccl().ok ? CUMICRO_OK : cmh::fail(CUMICRO_E_NODEVICE, "NCCL is not loaded in this process and libnccl.so.2 was not found");
}

}  // namespace

extern "C" {

int64_t cumicro_reduce_diagnostics_scratch_bytes(int ncols) { return (int64_t)sizeof(double) * kRedBlocks * (ncols < 1 ? 1 : ncols) + 16; }

int cumicro_reduce_diagnostics_f64(int64_t n, const double* weight, const double* const* cols, int ncols, double* out, void* scratch,
                                   int64_t scratch_bytes, void* stream) {
    return reduce_impl<double>(n, weight, cols, ncols, out, scratch, scratch_bytes, stream);
}
int cumicro_reduce_diagnostics_f32(int64_t n, const float* weight, const float* const* cols, int ncols, double* out, void* scratch,
                                   int64_t scratch_bytes, void* stream) {
    return reduce_impl<float>(n, weight, cols, ncols, out, scratch, scratch_bytes, stream);
}

int cumicro_nccl_allreduce_f64(void* comm, double* buf, int64_t count, void* stream) {
    if (comm == nullptr || buf == nullptr) return cmh::fail(CUMICRO_E_NULL, "comm / buf is NULL");
    if (count < 0) return cmh::fail(CUMICRO_E_SIZE, "count = %lld is negative", (long long)count);
    int st = need_nccl();
    if (st) return st;
    if (count == 0) return CUMICRO_OK;
    return nccl_status(nccl().all_reduce(buf, buf, (size_t)count, 8 /* ncclDouble */, 0 /* ncclSum */, comm, (cudaStream_t)stream), "ncclAllReduce");
}
int cumicro_nccl_unique_id(void* id128) {
    if (id128 == nullptr) return cmh::fail(CUMICRO_E_NULL, "id128 is NULL");
    int st = need_nccl();
    if (st) return st;
    return nccl_status(nccl().unique_id(static_cast<NcclId*>(id128)), "ncclGetUniqueId");
}
int cumicro_nccl_comm_init_rank(void** comm, int nranks, const void* id128, int rank) {
    if (comm == nullptr || id128 == nullptr) return cmh::fail(CUMICRO_E_NULL, "comm / id128 is NULL");
    if (nranks < 1 || rank < 0 || rank >= nranks) return cmh::fail(CUMICRO_E_ARG, "rank %d of %d", rank, nranks);
    int st = need_nccl();
    if (st) return st;
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    return nccl_status(nccl().init_rank(comm, nranks, id, rank), "ncclCommInitRank");
}
int cumicro_nccl_comm_destroy(void* comm) {
    if (comm == nullptr) return CUMICRO_OK;
    int st = need_nccl();
    if (st) return st;
    return nccl_status(nccl().destroy(comm), "ncclCommDestroy");
}

}  // extern "C"
