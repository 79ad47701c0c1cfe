// cm_host.cu — host-side plumbing of libcumicro.so: error reporting, launch
// accounting, device queries.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include <map>
#include <vector>
#include <utility>

#include "cm_hostpipe.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace

namespace cmh {

void release_tables();   // kernels_2m.cu: the cached ventilation tables

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_status(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return CUMICRO_OK;
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int num_sms() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return cached;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) cached = v;
        cached_dev = dev;
    }
    return cached;
}

namespace {
struct Workspace {
    void* ptr = nullptr;
    size_t bytes = 0;
};
struct ThreadPipeState {
    std::map<std::pair<int, int>, Workspace> ws;          // (device, slot) -> staging buffer
    std::map<std::pair<int, int>, cudaStream_t> streams;  // (device, slot) -> stream
    // no destructor: at thread/process teardown the CUDA context may already be gone;
    // cumicro_release_workspace() frees explicitly.
};
thread_local ThreadPipeState g_pipe;
}  // namespace

int workspace(int slot, size_t bytes, void** ptr) {
    int dev = 0;
    int rc = cuda_status(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    Workspace& w = g_pipe.ws[{dev, slot}];
    if (w.bytes < bytes) {
        if (w.ptr) cudaFree(w.ptr);
        w.ptr = nullptr;
        w.bytes = 0;
        rc = cuda_status(cudaMalloc(&w.ptr, bytes), "cudaMalloc (host pipeline staging)");
        if (rc) return rc;
        w.bytes = bytes;
    }
    *ptr = w.ptr;
    return CUMICRO_OK;
}

int slot_stream(int slot, cudaStream_t* s) {
    int dev = 0;
    int rc = cuda_status(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    auto key = std::make_pair(dev, slot);
    auto it = g_pipe.streams.find(key);
    if (it == g_pipe.streams.end()) {
        cudaStream_t st;
        rc = cuda_status(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate");
        if (rc) return rc;
        it = g_pipe.streams.emplace(key, st).first;
    }
    *s = it->second;
    return CUMICRO_OK;
}

// Scratch keyed by (device, stream): two calls in flight on different streams never share it, and a buffer that has to grow is
// retired (freed by release_workspaces()), not freed under a kernel that may still use it.
namespace {
thread_local std::map<std::pair<int, cudaStream_t>, Workspace> g_stream_ws;
thread_local std::vector<void*> g_retired;
}  // namespace
int stream_workspace(cudaStream_t s, size_t bytes, void** ptr) {
    int dev = 0;
    int rc = cuda_status(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    Workspace& w = g_stream_ws[{dev, s}];
    if (w.bytes < bytes) {
        if (w.ptr) g_retired.push_back(w.ptr);
        w.ptr = nullptr;
        w.bytes = 0;
        rc = cuda_status(cudaMalloc(&w.ptr, bytes), "cudaMalloc (per-stream scratch)");
        if (rc) return rc;
        w.bytes = bytes;
    }
    *ptr = w.ptr;
    return CUMICRO_OK;
}

void release_workspaces() {
    for (auto& kv : g_pipe.ws)
        if (kv.second.ptr) cudaFree(kv.second.ptr);
    g_pipe.ws.clear();
    for (auto& kv : g_stream_ws)
        if (kv.second.ptr) cudaFree(kv.second.ptr);
    g_stream_ws.clear();
    for (void* p : g_retired) cudaFree(p);
    g_retired.clear();
}

}  // namespace cmh

extern "C" {

int cumicro_version(void) { return CUMICRO_VERSION; }
const char* cumicro_last_error(void) { return g_err; }
int64_t cumicro_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void cumicro_release_workspace(void) { cmh::release_workspaces(); cmh::release_tables(); }

}  // extern "C"
