// cm_tile2m.cuh — the launch shape of the headline kernel (fused 2-moment warm-rain tendencies, cm_sb2006_fast.cuh).
//
// The body is issue-slot bound together with the FP64 pipe: a point costs ~310 FP64-pipe instructions, each
// holding the pipe for two cycles, so every other issue slot is all the rest of the kernel may use.  This shape removes the
// per-thread instructions that are not arithmetic:
//   * block-uniform tile loop (tile = BLOCK consecutive points, tiles dealt round-robin to the persistent blocks): the input
//     tiles are fetched by ONE thread with 1-D bulk asynchronous copies (cp.async.bulk global -> shared, completion on an
//     mbarrier, two tiles in flight) instead of 7 cp.async + 14 address instructions per thread and point;
//   * the ~70 host-derived constants of the body are NOT hoisted out of the loop: 70 loop-invariant doubles need 140 uniform
//     registers, ptxas has ~80 and spills the rest through vector registers and local memory (measured: 16 STL/LDL + 34 R2UR
//     per point).  They are read at their point of use through an address that depends on the tile counter (an opaque zero),
//     i.e. one uniform constant-bank load per use and no register held;
//   * stores are streaming 64/32-bit st.global.cs, one per output column.
// Columns must be 16-byte aligned for the bulk copies (any cudaMalloc'd or torch buffer is); otherwise, and for a partial last
// tile, the same kernel loads with guarded scalar loads (identical results).
#pragma once
#include "cm_launch.cuh"
#include "cm_sb2006_fast.cuh"

// Device copy of the verified ventilation table of a parameter block (cm_sb2006_fast.cuh), cached per device and block content;
// fills k.tab_inv_h / tab_u0.  nullptr if the table cannot be used (verification failed, allocation failed, stream capture in progress):
// the caller then runs the closed form.  The first call for a new parameter block builds the table on the host (~1 ms) and uploads
// 9.4 KB with a blocking copy; later calls are a lookup.  Defined in kernels_2m.cu.
namespace cmh { const double* w2_table(const cumicro_params_2m_warm_f64& p, cm::W2K& k); }

namespace cm {

template <class FT, int NIN> struct Tile2MArgs {
    W2K k;
    const double* tab;
    const FT* in[NIN];
    FT* out[4];
    int64_t n;
    int bulk_ok;   // every input column is 16-byte aligned: full tiles arrive by bulk copies (otherwise by guarded scalar loads, same bits)
};


// PPT points per thread: the PPT bodies of one thread are independent and share every constant load.
template <class FT, int NIN, int LIM, int BLOCK, int MINB, bool ALL_OUT, int PPT, bool TAB>
__global__ void __launch_bounds__(BLOCK, MINB) warm2m_tile_kernel(const __grid_constant__ Tile2MArgs<FT, NIN> a) {
    constexpr int TILE = BLOCK * PPT;
    constexpr int NWARP = BLOCK / 32;
    __shared__ __align__(128) FT stage[2][NIN][TILE];
    __shared__ __align__(16) double tab_s[TAB ? kTabDoubles : 2];
    __shared__ __align__(8) unsigned long long full[2];
    if (TAB) {
        for (int i = threadIdx.x; i < kTabDoubles / 2; i += BLOCK)
            reinterpret_cast<double2*>(tab_s)[i] = __ldg(reinterpret_cast<const double2*>(a.tab) + i);
    }
    math_tables_init<BLOCK, false>();
    math_tables_init_log2<BLOCK>();
    const int tid = threadIdx.x;
    const unsigned n_tiles = (unsigned)((a.n + TILE - 1) / TILE);
    const unsigned n_full = a.bulk_ok ? (unsigned)(a.n / TILE) : 0u;   // tiles [0, n_full) are complete and fetched by bulk copies
    constexpr unsigned kTileBytes = TILE * sizeof(FT);
    // the leader lane of warp w fetches columns w, w + NWARP, ...: the copy-issue instructions are spread over the SM sub-partitions
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction
    const bool issuer = warp < NIN;
    constexpr int kIssuers = NWARP < NIN ? NWARP : NIN;
    if (tid == 0) {
        tile::mbar_init(&full[0], kIssuers);
        tile::mbar_init(&full[1], kIssuers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](unsigned t, int buf) {   // converged issuer warps; one elected lane does the work
        if (t < n_full && tile::elect_one()) {
            const int ncol = (NIN - warp + NWARP - 1) / NWARP;
            tile::mbar_expect_tx(&full[buf], ncol * kTileBytes);
#pragma unroll 1
            for (int c = warp; c < NIN; c += NWARP) tile::bulk_g2s(&stage[buf][c][0], a.in[c] + (size_t)t * TILE, kTileBytes, &full[buf]);
        }
    };
    unsigned t = blockIdx.x;
    if (issuer) {   // two tiles in flight
        issue(t, 0);
        if (t + gridDim.x < n_tiles) issue(t + gridDim.x, 1);
    }
    unsigned j = 0;
    for (; t < n_tiles; t += gridDim.x, ++j) {
        const int buf = j & 1;
        const int64_t it0 = (int64_t)t * TILE + tid;
        double x[PPT][NIN];
        if (t < n_full) {
            tile::mbar_wait(&full[buf], (j >> 1) & 1);
#pragma unroll
            for (int p = 0; p < PPT; ++p)
#pragma unroll
                for (int c = 0; c < NIN; ++c) x[p][c] = (double)stage[buf][c][p * BLOCK + tid];
        } else {
#pragma unroll
            for (int p = 0; p < PPT; ++p)
#pragma unroll
                for (int c = 0; c < NIN; ++c) x[p][c] = (it0 + p * BLOCK < a.n) ? (double)__ldg(a.in[c] + it0 + p * BLOCK) : 1.0;
        }
        __syncthreads();   // every thread has read stage[buf]: it may be refilled
        if (issuer) {
            const unsigned t2 = t + 2 * gridDim.x;
            if (t2 < n_tiles) issue(t2, buf);
        }
        // The ~70 constants are read from the constant bank where they are used (LDCU.128 / LDC.64 per use): with the block barrier and
        // the asynchronous copies inside the loop ptxas does not hoist them (hoisted, they need 140 uniform registers; ptxas has ~80 and
        // spilled the rest through vector registers and local memory in the cp.async shape: 16 STL/LDL + 34 R2UR per point).
        const W2K& k = a.k;
        double y[PPT][4];
#pragma unroll
        for (int p = 0; p < PPT; ++p)
            warm2m_fast<LIM, TAB>(k, x[p][0], x[p][1], x[p][2], x[p][3], x[p][4], x[p][5], x[p][6], (NIN == 8) ? clamp0_(x[p][NIN - 1]) : 0.0,
                                  NIN == 8, y[p], tab_s);
#pragma unroll
        for (int p = 0; p < PPT; ++p)
            if (it0 + p * BLOCK < a.n) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (ALL_OUT || a.out[c]) __stcs(a.out[c] + it0 + p * BLOCK, (FT)y[p][c]);
            }
    }
}

// Enqueue the tile kernel.  `tab` = device ventilation table of this block (cmh::w2_table) or nullptr (closed form).
// Columns that are not 16-byte aligned are read with scalar loads by the same kernel (identical results).
// Launch shape (tools/tune_2m.py, 2^24 points, ms): 128x5 0.440, 128x6 0.407, 128x7 0.416, 128x8 0.423, 256x3 0.467 (16 waves);
// 128x7 with 4 / 8 / 16 / 32 / 64 waves of blocks: 0.414 / 0.411 / 0.416 / 0.422 / 0.463.
template <class FT, int NIN, int LIM, int BLOCK = 128, int MINB = 6, int PPT = 1>
int launch_warm2m_tile(const W2K& k, const double* tab, int64_t n, const FT* const (&in)[NIN], FT* const (&out)[4], cudaStream_t stream,
                       const char* what, int waves = 8) {
    if (n == 0) return CUMICRO_OK;
    Tile2MArgs<FT, NIN> a;
    a.bulk_ok = 1;
    for (int c = 0; c < NIN; ++c)
        if (!cmh::aligned16(in[c])) a.bulk_ok = 0;
    a.k = k;
    a.tab = tab;
    a.n = n;
    bool all_out = true;
    for (int c = 0; c < NIN; ++c) a.in[c] = in[c];
    for (int c = 0; c < 4; ++c) { a.out[c] = out[c]; all_out = all_out && out[c] != nullptr; }
    constexpr bool kCanTab = LIM == 1;
    const bool use_tab = kCanTab && tab != nullptr;
    void (*kern)(Tile2MArgs<FT, NIN>);
    if (use_tab) kern = all_out ? warm2m_tile_kernel<FT, NIN, LIM, BLOCK, MINB, true, PPT, kCanTab> : warm2m_tile_kernel<FT, NIN, LIM, BLOCK, MINB, false, PPT, kCanTab>;
    else kern = all_out ? warm2m_tile_kernel<FT, NIN, LIM, BLOCK, MINB, true, PPT, false> : warm2m_tile_kernel<FT, NIN, LIM, BLOCK, MINB, false, PPT, false>;
    static bool carveout_set[4] = {false, false, false, false};
    const int slot = (use_tab ? 2 : 0) + (all_out ? 1 : 0);
    if (!carveout_set[slot]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carveout_set[slot] = true;
    }
#ifdef CUMICRO_TUNING
    { const char* wv0 = getenv("CUMICRO_WAVES"); if (wv0) waves = atoi(wv0); }
#endif
    constexpr int TILE = BLOCK * PPT;
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    const int blocks = (int)std::min<int64_t>(n_tiles, (int64_t)cmh::num_sms() * MINB * waves);
    kern<<<blocks, BLOCK, 0, stream>>>(a);
    cmh::count_launch();
    return cmh::cuda_status(cudaGetLastError(), what);
}

}  // namespace cm
