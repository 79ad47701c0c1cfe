// cm_p2p.cuh — the diagnostics all-reduce as peer-memory stores inside the kernel that produces the sums.
//
// The only exchange step of the path sums CUMICRO_NDIAG doubles over the slabs' GPUs.  As a library call it is a second kernel
// launch plus NCCL's protocol (15 us at 2 GPUs, 27 us at 8, bench.py config5); here it is the tail of the kernel that finishes the
// slab's own sums: every rank owns a WINDOW in its device memory with one slot per rank and call parity, writes its `count` doubles
// straight into its slot of every peer's window over NVLink / NVSwitch (ordinary stores to peer-mapped memory, then a release
// store of the call number), waits until its own window holds the call number from every rank, and adds the slots in RANK order
// — so all ranks obtain bit-identical sums without a second pass.
//
// Ordering argument for the two parities: rank A writes call s + 2 into the slots call s used.  It does so only after finishing
// call s + 1, i.e. after every rank's call-(s + 1) contribution has arrived, and a rank sends its call-(s + 1) contribution only
// after it has finished READING call s.  The call number lives in the owner's window (device side), so a captured CUDA graph that
// replays the kernel stays in step with its peers.  A wait gives up after `timeout_ns` (a peer that never calls): the result is
// NaN and the window's error word holds the call number (cumicro_p2p_window_status).
#pragma once
#include <cstdint>

namespace cm {

constexpr int kP2PMaxRanks = 16;   // one NVLink / NVSwitch domain (8 GPUs on the B200 boards of this pool)
constexpr int kP2PMaxCount = 16;

struct alignas(128) P2PSlot {
    double v[kP2PMaxCount];
    unsigned long long seq;
    unsigned long long pad[15];
};
struct P2PWindowMem {   // in the owner's device memory; peer r writes slots[*][r]
    P2PSlot slots[2][kP2PMaxRanks];
    unsigned long long counter;   // calls started by the owner
    unsigned long long error;     // 0, or the call number whose wait timed out
};
struct P2PDev {   // kernel argument (by value)
    P2PWindowMem* win[kP2PMaxRanks];   // win[rank] = the local window, the others are peer mappings
    int rank, nranks;
    unsigned long long timeout_ns;
};

#ifdef __CUDACC__
__device__ __forceinline__ void p2p_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long p2p_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void p2p_st_sys(double* p, double v) { asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double p2p_ld_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long p2p_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// In-place sum of buf[0..count) over the ranks; called by EVERY thread of ONE block (blockDim.x >= max(nranks, count)), buf
// already visible to the block (written before a __syncthreads()).  The result is in buf after the function returns (block-wide).
__device__ inline void p2p_allreduce_block(const P2PDev& d, double* buf, int count) {
    __shared__ unsigned long long s_seq;
    __shared__ int s_bad;
    const int t = threadIdx.x;
    P2PWindowMem* mine = d.win[d.rank];
    if (t == 0) {
        s_seq = ++mine->counter;   // only the owner's kernels touch the counter, and they are ordered on its stream
        s_bad = 0;
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int par = (int)(seq & 1ull);
    if (t < d.nranks) {   // thread t serves rank t: send, then wait for its contribution
        P2PSlot* dst = &d.win[t]->slots[par][d.rank];
        for (int k = 0; k < count; ++k) p2p_st_sys(&dst->v[k], buf[k]);
        __threadfence_system();
        p2p_st_release_sys(&dst->seq, seq);
        const P2PSlot* src = &mine->slots[par][t];
        const unsigned long long t0 = p2p_globaltimer();
        unsigned spins = 0;
        while (p2p_ld_acquire_sys(&src->seq) < seq) {
            if ((++spins & 1023u) == 0 && p2p_globaltimer() - t0 > d.timeout_ns) {
                atomicExch(&s_bad, 1);
                break;
            }
        }
    }
    __syncthreads();
    const int bad = s_bad;
    if (t < count) {
        double s = 0.0;
        for (int r = 0; r < d.nranks; ++r) s += p2p_ld_sys(&mine->slots[par][r].v[t]);   // rank order: the same bits on every rank
        buf[t] = bad ? __longlong_as_double(0x7ff8000000000000LL) : s;
    }
    if (t == 0 && bad) mine->error = seq;
    __syncthreads();
}
#endif  // __CUDACC__

}  // namespace cm

namespace cmh {
// Kernel-side view of a connected window (cumicro_p2p_window_*; cm_collective.cu).  Non-zero return: the window is NULL / not connected.
int p2p_dev(void* win, cm::P2PDev* out);
}  // namespace cmh
