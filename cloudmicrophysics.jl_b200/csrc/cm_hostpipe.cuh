// cm_hostpipe.cuh — host-buffer form of a pointwise family (the `_host` C-ABI entry
// points): a chunked H2D -> kernel -> D2H pipeline.
//
// NSLOT slots, each with its own stream and a device staging area of
// (NIN + NOUT) * chunk elements.  Slot s processes chunks s, s+NSLOT, ...; inside a
// slot the stream orders copy-in -> kernel -> copy-out, and the slots overlap each
// other, so the two DMA engines (one per direction) and the SMs all stay busy.
// With pinned host memory the copies are truly asynchronous; pageable memory still
// works (the driver stages it) but serialises more.  The staging areas are cached
// per host thread and device (cmh::workspace) so steady-state calls allocate nothing.
#pragma once
#include <algorithm>

#include "cm_types.cuh"

namespace cmh {
constexpr int kPipeSlots = 3;
// Per-thread, per-device cached device buffer for pipeline slot `slot` (grown on demand).
int workspace(int slot, size_t bytes, void** ptr);
// Per-thread non-blocking stream of slot `slot` on the current device.
int slot_stream(int slot, cudaStream_t* s);
void release_workspaces();
// Per-thread device scratch of the calling (device, stream): never shared between streams, never freed while it may be in use.
int stream_workspace(cudaStream_t s, size_t bytes, void** ptr);
}  // namespace cmh

namespace cm {

template <class FT, int NIN, int NOUT, class LaunchFn>
int host_pipeline(int64_t n, const FT* const (&in)[NIN], FT* const (&out)[NOUT], int64_t chunk, LaunchFn launch) {
    if (n == 0) return CUMICRO_OK;
    if (chunk <= 0) chunk = int64_t(1) << 20;
    chunk = (chunk + 63) & ~int64_t(63);  // keep every staged column 16-byte aligned
    if (chunk > n) chunk = (n + 63) & ~int64_t(63);
    const size_t col_bytes = sizeof(FT) * (size_t)chunk;
    const int64_t n_chunks = (n + chunk - 1) / chunk;
    const int n_slots = (int)std::min<int64_t>(cmh::kPipeSlots, n_chunks);
    cudaStream_t st[cmh::kPipeSlots];
    char* ws[cmh::kPipeSlots];
    for (int s = 0; s < n_slots; ++s) {
        int rc = cmh::slot_stream(s, &st[s]);
        if (rc) return rc;
        void* p = nullptr;
        if ((rc = cmh::workspace(s, col_bytes * (NIN + NOUT), &p))) return rc;
        ws[s] = static_cast<char*>(p);
    }
    int rc = CUMICRO_OK;
    for (int64_t c = 0; c < n_chunks && rc == CUMICRO_OK; ++c) {
        const int s = (int)(c % n_slots);
        const int64_t i0 = c * chunk;
        const int64_t m = std::min<int64_t>(chunk, n - i0);
        const FT* din[NIN];
        FT* dout[NOUT];
        for (int k = 0; k < NIN; ++k) {
            FT* d = reinterpret_cast<FT*>(ws[s] + col_bytes * k);
            rc = cmh::cuda_status(cudaMemcpyAsync(d, in[k] + i0, sizeof(FT) * (size_t)m, cudaMemcpyHostToDevice, st[s]),
                                  "host pipeline H2D");
            if (rc) break;
            din[k] = d;
        }
        if (rc) break;
        for (int k = 0; k < NOUT; ++k) dout[k] = out[k] ? reinterpret_cast<FT*>(ws[s] + col_bytes * (NIN + k)) : nullptr;
        rc = launch(m, din, dout, st[s]);
        if (rc) break;
        for (int k = 0; k < NOUT; ++k) {
            if (!out[k]) continue;
            rc = cmh::cuda_status(
                cudaMemcpyAsync(out[k] + i0, dout[k], sizeof(FT) * (size_t)m, cudaMemcpyDeviceToHost, st[s]),
                "host pipeline D2H");
            if (rc) break;
        }
    }
    for (int s = 0; s < n_slots; ++s) {
        int rc2 = cmh::cuda_status(cudaStreamSynchronize(st[s]), "host pipeline sync");
        if (rc == CUMICRO_OK) rc = rc2;
    }
    return rc;
}

}  // namespace cm
