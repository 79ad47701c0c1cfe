// cm_collective.cu — the only exchange step of the path: domain diagnostics of the column slabs.
//
//   cumicro_reduce_diagnostics_*  one slab's sums  out[k] = Σ_i w[i] cols[k][i]  (w = rho or NULL = 1), Float64 accumulation in a
//                                 fixed order (bit-reproducible, independent of the launch shape): the diagnostics of tendencies
//                                 that were computed by the non-fused entry points (the fused config-5 kernel reduces in-kernel).
//   cumicro_nccl_allreduce_f64    sum of `count` doubles over the ranks of the caller's ncclComm_t, in place, on the caller's stream.
//   cumicro_nccl_{unique_id,comm_init_rank,comm_destroy}   for hosts without an NCCL binding of their own.
//   cumicro_p2p_window_* / cumicro_p2p_allreduce_f64   the same sum as peer-memory stores in ONE kernel (cm_p2p.cuh): no library, no
//                                 second launch; the fused config-5 entry point takes the window and reduces in its finish kernel.
// libcumicro.so has no link-time dependency on NCCL: the symbols are resolved at first use from the NCCL already loaded in the process
// (torch / NCCL.jl / MPI stack) or from libnccl.so.2 on the loader path.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "cm_hostpipe.cuh"
#include "cm_p2p.cuh"

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

// ---- peer-memory window (cm_p2p.cuh) --------------------------------------------------------------------------------------
struct P2PWindow {
    int rank = 0, nranks = 1, device = 0;
    cm::P2PWindowMem* local = nullptr;
    cm::P2PWindowMem* peer[cm::kP2PMaxRanks] = {};
    bool opened[cm::kP2PMaxRanks] = {};
    bool connected = false;
    unsigned long long timeout_ns = 10ull * 1000000000ull;
};

__global__ void __launch_bounds__(32) p2p_allreduce_kernel(const __grid_constant__ cm::P2PDev d, double* buf, int count) {
    cm::p2p_allreduce_block(d, buf, count);
}

}  // namespace

namespace cmh {
int p2p_dev(void* win, cm::P2PDev* out) {
    auto* w = static_cast<P2PWindow*>(win);
    if (w == nullptr || out == nullptr) return fail(CUMICRO_E_NULL, "p2p window is NULL");
    if (!w->connected) return fail(CUMICRO_E_ARG, "p2p window of rank %d is not connected (cumicro_p2p_window_connect)", w->rank);
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != w->device) return fail(CUMICRO_E_ARG, "p2p window belongs to device %d, the current device is %d", w->device, dev);
    for (int r = 0; r < cm::kP2PMaxRanks; ++r) out->win[r] = r < w->nranks ? w->peer[r] : nullptr;
    out->rank = w->rank;
    out->nranks = w->nranks;
    out->timeout_ns = w->timeout_ns;
    return CUMICRO_OK;
}
}  // namespace cmh

namespace {

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

int cumicro_p2p_window_create(int rank, int nranks, void** win) {
    if (win == nullptr) return cmh::fail(CUMICRO_E_NULL, "win is NULL");
    *win = nullptr;
    if (nranks < 1 || nranks > cm::kP2PMaxRanks || rank < 0 || rank >= nranks)
        return cmh::fail(CUMICRO_E_ARG, "rank %d of %d (1 <= nranks <= %d)", rank, nranks, cm::kP2PMaxRanks);
    auto* w = new P2PWindow();
    w->rank = rank;
    w->nranks = nranks;
    int st = cmh::cuda_status(cudaGetDevice(&w->device), "cudaGetDevice");
    // cudaMalloc (not a stream-ordered pool): only such allocations can be exported with cudaIpcGetMemHandle
    if (!st) st = cmh::cuda_status(cudaMalloc(reinterpret_cast<void**>(&w->local), sizeof(cm::P2PWindowMem)), "cudaMalloc (p2p window)");
    if (!st) st = cmh::cuda_status(cudaMemset(w->local, 0, sizeof(cm::P2PWindowMem)), "cudaMemset (p2p window)");
    if (!st) st = cmh::cuda_status(cudaDeviceSynchronize(), "p2p window create");   // zero before any peer can hold the handle
    if (st) {
        if (w->local) cudaFree(w->local);
        delete w;
        return st;
    }
    w->peer[rank] = w->local;
    w->connected = nranks == 1;
    *win = w;
    return CUMICRO_OK;
}
int cumicro_p2p_window_handle(void* win, void* handle64) {
    auto* w = static_cast<P2PWindow*>(win);
    if (w == nullptr || handle64 == nullptr) return cmh::fail(CUMICRO_E_NULL, "win / handle64 is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == CUMICRO_P2P_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    int st = cmh::cuda_status(cudaIpcGetMemHandle(&h, w->local), "cudaIpcGetMemHandle");
    if (st) return st;
    memcpy(handle64, &h, sizeof(h));
    return CUMICRO_OK;
}
int cumicro_p2p_window_connect(void* win, const void* handles) {
    auto* w = static_cast<P2PWindow*>(win);
    if (w == nullptr) return cmh::fail(CUMICRO_E_NULL, "win is NULL");
    if (w->connected) return CUMICRO_OK;
    if (handles == nullptr) return cmh::fail(CUMICRO_E_NULL, "handles is NULL");
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != w->device) return cmh::fail(CUMICRO_E_ARG, "p2p window belongs to device %d, the current device is %d", w->device, dev);
    for (int r = 0; r < w->nranks; ++r) {
        if (r == w->rank || w->opened[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        int st = cmh::cuda_status(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle (peer window; the ranks must be separate processes on one NVLink node)");
        if (st) return st;
        w->peer[r] = static_cast<cm::P2PWindowMem*>(p);
        w->opened[r] = true;
    }
    w->connected = true;
    return CUMICRO_OK;
}
int cumicro_p2p_window_set_timeout(void* win, double seconds) {
    auto* w = static_cast<P2PWindow*>(win);
    if (w == nullptr) return cmh::fail(CUMICRO_E_NULL, "win is NULL");
    if (!(seconds > 0.0) || seconds > 3600.0) return cmh::fail(CUMICRO_E_ARG, "timeout = %g s (0 < t <= 3600)", seconds);
    w->timeout_ns = (unsigned long long)(seconds * 1e9);
    return CUMICRO_OK;
}
int cumicro_p2p_window_status(void* win, int64_t* calls, int64_t* timed_out_call) {
    auto* w = static_cast<P2PWindow*>(win);
    if (w == nullptr) return cmh::fail(CUMICRO_E_NULL, "win is NULL");
    unsigned long long v[2] = {0, 0};
    int st = cmh::cuda_status(cudaMemcpy(v, &w->local->counter, sizeof(v), cudaMemcpyDeviceToHost), "p2p window status");
    if (st) return st;
    if (calls) *calls = (int64_t)v[0];
    if (timed_out_call) *timed_out_call = (int64_t)v[1];
    return CUMICRO_OK;
}
int cumicro_p2p_allreduce_f64(void* win, double* buf, int count, void* stream) {
    if (buf == nullptr) return cmh::fail(CUMICRO_E_NULL, "buf is NULL");
    if (count < 0 || count > cm::kP2PMaxCount) return cmh::fail(CUMICRO_E_SIZE, "count = %d (0..%d)", count, cm::kP2PMaxCount);
    cm::P2PDev d;
    int st = cmh::p2p_dev(win, &d);
    if (st) return st;
    if (count == 0) return CUMICRO_OK;
    p2p_allreduce_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d, buf, count);
    cmh::count_launch();
    return cmh::cuda_status(cudaGetLastError(), "p2p_allreduce launch");
}
int cumicro_p2p_window_destroy(void* win) {
    auto* w = static_cast<P2PWindow*>(win);
    if (w == nullptr) return CUMICRO_OK;
    for (int r = 0; r < w->nranks; ++r)
        if (w->opened[r]) cudaIpcCloseMemHandle(w->peer[r]);
    if (w->local) cudaFree(w->local);
    delete w;
    return CUMICRO_OK;
}

}  // extern "C"
