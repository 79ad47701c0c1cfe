// kernels_p3.cu — P3 ice scheme kernels and their C-ABI entry points (include/cumicro.h):
//   cumicro_p3_rates_*     stand-alone P3 terminal velocities + process rates (BASELINE config 4)
//   cumicro_bmt2m_p3_*     BMT.bulk_microphysics_tendencies(::Microphysics2Moment, mp{WR,<:P3IceParams}, ...)  BMT:898-1083
//   cumicro_termvel_p3_*   P3.ice_terminal_velocity_{number,mass}_weighted_from_prognostic   P3_terminal_velocity.jl:135-173
//   cumicro_p3_logl_*      P3.get_distribution_logλ_from_prognostic                          P3_size_distribution.jl:284-334
//
// Kernel shape (cm_p3.cuh): a warp owns a tile of 32 consecutive points.  Lanes load their own
// point (coalesced), the cheap pointwise parts (warm rain, nucleation, deposition, number
// adjustment) run one point per lane, and the quadrature-based ice processes of every
// ice-bearing point of the tile are evaluated by the whole warp, one point after the other.
// Tiles are dealt to warps round-robin so that ice-free and ice-bearing regions of the grid mix.
#include <algorithm>
#include <cmath>

#include "cm_launch.cuh"
#include "cm_p3.cuh"

namespace {

using namespace cm;

// Launch shape: ONE block of 896 threads per SM (28 warps at 72 registers) and ONE block barrier per point (CUMICRO_P3_SYNC = 1):
// the kernel is instruction-fetch bound — a point's pass walks ~90 KB of code against a 6 KB L0 / 32 KB L1.5
// instruction cache — and warps that start every point together walk the same loops at the same time and share the lines.  The
// more of an SM's warps share one barrier, the faster: 2^20 points, process rates (tools/tune_p3.py, round 2, after the quantile
// solves moved ahead of the passes): 256x4 97.8 ms | 320x3 72.7 | 512x2 65.8 | 768x1 63.5 | 896x1 57.7 | 1024x1 (64 registers)
// 60.2; a barrier every 2nd / 4th / no point (512x2): 77.9 / 86.6 / 90.8; barriers at every phase as well: 512x2 63.1, 1024x1 63.7.
#ifndef CUMICRO_P3_BLOCK
#define CUMICRO_P3_BLOCK 896
#define CUMICRO_P3_MINB 1
#endif
#ifndef CUMICRO_P3_SYNC
#define CUMICRO_P3_SYNC 1
#endif
#ifndef CUMICRO_P3_SYNC_EVERY
#define CUMICRO_P3_SYNC_EVERY 1
#endif
#ifndef CUMICRO_P3L_BLOCK
#define CUMICRO_P3L_BLOCK 896   /* the stand-alone logλ solve, 2^22 points: 128x3 19.9 ms, 128x6 17.9, 512x2 15.2, 640x1 16.0, 768x1 15.0, 896x1 14.7, 1024x1 14.8 */
#define CUMICRO_P3L_MINB 1     /* (one block per SM: the warps walk the ~70 KB solve together, like the main kernel) */
#endif
constexpr int BLOCK = CUMICRO_P3_BLOCK;
constexpr int MINB = CUMICRO_P3_MINB;
enum { MODE_RATES = 0, MODE_BMT = 1, MODE_VEL = 2 };
constexpr int NIN_MAX = 13, NOUT_MAX = 12;
constexpr int kSlot = 17;   // doubles per point in the owner <-> evaluating-warp exchange (11 inputs + 6 quantiles out, 12 rates + F_rim, rho_rim back)

template <class FT> struct P3Args {
    cumicro_params_p3_f64 p;
    ThermoK<double> tk;
    SB2006K<double> sk;
    P3K k;
    const FT* in[NIN_MAX];
    FT* out[NOUT_MAX];
    int64_t n;
    int want;
    int solve_logl;   // the logλ column is NULL: solve it in the kernel, one listed point per thread, before the quantile phase
};

// P3.get_distribution_logλ_from_prognostic of one point (defined below, with cumicro_p3_logl_*: the same code, the same bits)
__device__ double p3_solve_logl(const P3K& k, double L_ice, double N_ice, double L_rim, double B_rim, int iters);

// ---- pointwise parts of BMT:898-1083 -------------------------------------------------------------
// IN.INP_concentration_mean                                                        IN:250-253
CM_DEV double inp_log_mean(const P3K& k, double T) {
    const double Tc = fmin_(T - k.frost_T_freeze, 0.0);
    return 9.0 * log_full_(-k.frost_b * Tc / 10.0) - k.frost_log_a;
}

struct Pt {   // one grid point, clamped (BMT:911-930)
    double rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, q_rim, b_rim, logl, shift;
    double L_lcl, N_lcl, L_rai, N_rai, L_ice, N_ice, L_rim, B_rim;
};

// IN.liquid_freezing_rate (rain: IN:274-311, cloud PSD: IN:356-388), IN.immersion_limit_rate (IN:420-430) and
// IN.deposition_rate (IN:491-511) with n_active = n_ice (IN:526), as BMT:998-1075 calls them.
struct F23Rates { double rain_dn, rain_dq, cld_dn, cld_dq, cap_dn, dep_dn, dep_dq; };
CM_DEV F23Rates f23_rates(const cumicro_params_p3_f64& p, const ThermoK<double>& tk, const SB2006K<double>& sk, const P3K& k, const Pt& x) {
    F23Rates o;
    const double e = tk.eps;
    const double rho = x.rho, T = x.T;
    const TempState<double> ts = temp_state(tk, T);
    const double q_sat_ice = p_sat_ice(tk, ts) / (tk.R_v * rho * T);
    const double q_liq = x.q_lcl + x.q_rai;
    const double qv = q_vap(x.q_tot, q_liq, x.q_ice);
    const double inpc = exp_full_(inp_log_mean(k, T) + x.shift) / rho;
    {
        const double S_i = qv / q_sat_ice - 1.0;
        const bool cond = (T < k.frost_T_freeze - 15.0) && (S_i > 0.05);
        const double a = clamp0_(inpc - x.n_ice) / k.tau_act;
        o.dep_dn = cond ? a : 0.0;
        const double q_excess = clamp0_(qv - q_sat_ice);
        o.dep_dq = fmin_(k.m_nuc * o.dep_dn, q_excess / (2.0 * k.tau_act));
    }
    const double J_bigg = k.het_B * exp_full_(k.het_a * (tk.T_freeze - T));
    {
        const auto& pc = p.warm.sb.pdf_c;
        const double n = x.N_lcl / rho;
        const bool off = (x.N_lcl < e) || (x.q_lcl < e);
        const double safe_q = fmax_(x.q_lcl, e), safe_N = fmax_(x.N_lcl, e);
        const double logx = log_full_(rho * safe_q / safe_N);
        const double lB = -pc.mu_c * (logx + pc.loggamma_z1 - pc.loggamma_z2);
        const double loglam_c = off ? num<double>::inf() : lB + pc.mu_c * k.log_km;
        const double M3 = n * exp_full_(-3.0 / k.mu_cD * loglam_c) * k.cloud_M3_ratio;
        const double M6 = n * exp_full_(-6.0 / k.mu_cD * loglam_c) * k.cloud_M6_ratio;
        const bool cond = (n > e) && (x.q_lcl > e) && (T < tk.T_freeze - 4.0);
        o.cld_dn = cond ? J_bigg * k.V1 * M3 : 0.0;
        o.cld_dq = cond ? J_bigg * k.rho_w * (k.V1 * k.V1) * M6 : 0.0;
        o.cap_dn = (T >= k.frost_T_freeze) ? 0.0 : clamp0_(inpc - x.n_ice) / k.tau_act;
    }
    {
        const double n = x.N_rai / rho;
        const RainPDF<double> rp = pdf_rain_parameters<double>(p.warm.sb.pdf_r, sk.pi_rho_w, e, x.q_rai, rho, x.N_rai);
        const double Dr = rp.Dr_mean, D3 = Dr * Dr * Dr;
        const double M3 = n * 6.0 * D3, M6 = n * 720.0 * (D3 * D3);
        const bool cond = (n > e) && (x.q_rai > e) && (T < tk.T_freeze - 4.0);
        o.rain_dn = cond ? J_bigg * k.V1 * M3 : 0.0;
        o.rain_dq = cond ? J_bigg * k.rho_w * (k.V1 * k.V1) * M6 : 0.0;
    }
    return o;
}

CM_DEV void bmt2m_p3_assemble(const cumicro_params_p3_f64& p, const ThermoK<double>& tk, const SB2006K<double>& sk, const P3K& k,
                              const Pt& x, bool ice_on, const P3Rates& r, double F_rim, double rho_rim, double (&y)[9]) {
    const double e = tk.eps;
    const double rho = x.rho, T = x.T;
    const Warm2M<double> w = warm_rain_tendencies_2m<double>(p.warm, tk, sk, rho, T, x.q_tot, x.q_lcl, x.n_lcl, x.q_rai, x.n_rai, x.q_ice);
    double dq_lcl = w.dq_lcl_dt, dn_lcl = w.dn_lcl_dt, dq_rai = w.dq_rai_dt, dn_rai = w.dn_rai_dt;
    double dq_ice = 0.0, dn_ice = 0.0, dq_rim = 0.0, db_rim = 0.0;
    if (ice_on) {                                                                   // BMT:961-996
        dq_lcl += r.src[0];
        dq_rai += r.src[1];
        dn_lcl += r.src[2] / rho;
        dn_rai += r.src[3] / rho;
        dq_ice += r.src[5] / rho;
        dq_rim += r.src[4] / rho;
        db_rim += r.src[6] / rho;
        dn_ice -= r.agg_dN / rho;
        const double dq_m = r.melt_dL / rho, dn_m = r.melt_dN / rho;
        dq_rai += dq_m;
        dn_rai += dn_m;
        dq_ice -= dq_m;
        dn_ice -= dn_m;
        dq_rim -= dq_m * F_rim;
        db_rim -= (rho_rim > 0.0) ? dq_m * F_rim / rho_rim : 0.0;
    }
    // ---- F23 deposition nucleation, Bigg freezing of cloud drops capped by F23    BMT:998-1036
    const F23Rates f = f23_rates(p, tk, sk, k, x);
    const TempState<double> ts = temp_state(tk, T);
    const double q_sat_ice = p_sat_ice(tk, ts) / (tk.R_v * rho * T);
    const double q_liq = x.q_lcl + x.q_rai;
    const double qv = q_vap(x.q_tot, q_liq, x.q_ice);
    dn_ice += f.dep_dn;
    dq_ice += f.dep_dq;
    {
        const double dn_imm = fmin_(f.cld_dn, f.cap_dn);
        const double dq_imm = (f.cld_dn > 0.0) ? f.cld_dq * dn_imm / f.cld_dn : 0.0;
        dq_lcl -= dq_imm;
        dn_lcl -= dn_imm;
        dq_ice += dq_imm;
        dn_ice += dn_imm;
        dq_rim += dq_imm;
        db_rim += dq_imm / k.rho_i;
    }
    // ---- ice deposition / sublimation                                            BMT:1038-1054, NEQ:168-193
    {
        const double n_per_q = (x.q_ice > e) ? x.n_ice / x.q_ice : 0.0;
        const double Ls = latent_heat_sublim(tk, T);
        const double cp_air = cp_m(tk, x.q_tot, q_liq, x.q_ice + 0.0);
        const double dqsi_dT = q_sat_ice * (Ls / (tk.R_v * (T * T)) - 1.0 / T);
        const double Gam = 1.0 + Ls / cp_air * dqsi_dT;
        const double se = qv - q_sat_ice;
        const double timescale = k.subdep_tau * Gam;
        double tend = (se < 0.0) ? -fmin_(-se, clamp0_(x.q_ice)) / timescale : se / timescale;
        tend = ((T > tk.T_freeze) && (tend > 0.0)) ? 0.0 : tend;                       // NEQ.INP_limiter
        tend = (T > tk.T_freeze) ? fmin_(tend, 0.0) : tend;
        const double dn_d = (tend < 0.0) ? n_per_q * tend : 0.0;
        dq_ice += tend;
        dn_ice += dn_d;
        const double sub = fmin_(tend, 0.0);
        dq_rim += sub * F_rim;
        db_rim += (rho_rim > 0.0) ? sub * F_rim / rho_rim : 0.0;
    }
    // ---- ice number adjustment (τ = 100, x in [1e-12, 1e-5])                       BMT:1056-1064
    dn_ice += number_tendency_from_mass_limits<double>(e, 1.0 / 1e-12, 1.0 / 1e-5, 1.0 / 100.0, x.q_ice, x.n_ice);
    // ---- rain Bigg freezing                                                        BMT:1066-1075
    {
        const double rn = f.rain_dn, rq = f.rain_dq;
        dq_rai -= rq;
        dn_rai -= rn;
        dq_ice += rq;
        dn_ice += rn;
        dq_rim += rq;
        db_rim += rq / k.rho_i;
    }
    y[0] = dq_lcl; y[1] = dn_lcl; y[2] = dq_rai; y[3] = dn_rai; y[4] = dq_ice; y[5] = dn_ice; y[6] = dq_rim; y[7] = db_rim;
    y[8] = 0.0;   // dn_lcl_activation_dt: plumbed but zero in the reference (BMT:729, 1077-1078)
}

template <class FT, int MODE>
__global__ void __launch_bounds__(BLOCK, MINB) p3_tile_kernel(const __grid_constant__ P3Args<FT> a) {
    extern __shared__ double smem[];
    math_tables_init<BLOCK>();
    const int nq = a.k.n;
    double* qx = smem;
    double* qw = smem + nq;
    for (int i = threadIdx.x; i < nq; i += BLOCK) {
        qx[i] = a.p.quad.nodes[i];
        qw[i] = a.p.quad.weights[i];
    }
    __syncthreads();
    P3Scratch sc;
    constexpr int W = BLOCK / 32;
    sc.bind(smem + 2 * nq + (threadIdx.x >> 5) * P3Scratch::doubles(nq), nq);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double e = a.k.eps;
#if CUMICRO_P3_SYNC
    // ---- block-balanced form.  A round = BLOCK consecutive points, one per thread.  The points that need integrals are
    // compacted into a list in shared memory and dealt to the W warps round-robin, one point per warp per iteration, every
    // iteration starting with a block barrier (the SM's warps walk the same loops together: see the launch-shape note).
    // Inputs travel owner -> evaluating warp and the rates travel back through a 14-double slot per point.
    // the warp-uniform P3Point of the point a warp is evaluating lives in shared memory, ONE copy per warp: as a local it was 32
    // identical 280-byte copies per warp in local memory (72 registers cannot hold it), whose write-backs reached DRAM
    constexpr int kPointDoubles = (sizeof(P3Point) + 7) / 8;
    double* points = smem + 2 * nq + W * P3Scratch::doubles(nq);                     // [W][kPointDoubles]
    double* slots = points + W * kPointDoubles;                                      // [BLOCK][kSlot]
    unsigned short* idx = reinterpret_cast<unsigned short*>(slots + BLOCK * kSlot);   // [BLOCK] owners of the listed points
    unsigned char* wantv = reinterpret_cast<unsigned char*>(idx + BLOCK);             // [BLOCK]
    constexpr int kCats = 4;
    __shared__ int warp_cnt[kCats * W];
    for (int64_t base = (int64_t)blockIdx.x * BLOCK; base < a.n; base += (int64_t)gridDim.x * BLOCK) {
        const int64_t i = base + threadIdx.x;
        const bool valid = i < a.n;
        Pt x;
        auto ld = [&](int c) { return (valid && a.in[c]) ? (double)__ldg(a.in[c] + i) : 0.0; };
        if (MODE == MODE_VEL) {   // in: rho_a, L_ice, N_ice, L_rim, B_rim, logl (volumetric, as the reference's wrapper takes them)
            x.rho = ld(0); x.T = 273.15; x.q_tot = 0.0; x.q_lcl = x.n_lcl = x.q_rai = x.n_rai = 0.0;
            x.L_ice = ld(1); x.N_ice = ld(2); x.L_rim = ld(3); x.B_rim = ld(4); x.logl = ld(5); x.shift = 0.0;
            x.q_ice = x.n_ice = x.q_rim = x.b_rim = 0.0;
            x.L_lcl = x.N_lcl = x.L_rai = x.N_rai = 0.0;
        } else {                  // in: rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, q_rim, b_rim, logl [, inpc_log_shift]
            x.rho = clamp0_(ld(0)); x.T = ld(1); x.q_tot = clamp0_(ld(2)); x.q_lcl = clamp0_(ld(3)); x.n_lcl = clamp0_(ld(4));
            x.q_rai = clamp0_(ld(5)); x.n_rai = clamp0_(ld(6)); x.q_ice = clamp0_(ld(7)); x.n_ice = clamp0_(ld(8));
            x.q_rim = clamp0_(ld(9)); x.b_rim = clamp0_(ld(10)); x.logl = ld(11); x.shift = ld(12);
            x.L_lcl = x.q_lcl * x.rho; x.L_rai = x.q_rai * x.rho; x.N_lcl = x.n_lcl * x.rho; x.N_rai = x.n_rai * x.rho;
            x.L_ice = x.q_ice * x.rho; x.N_ice = x.n_ice * x.rho; x.L_rim = x.q_rim * x.rho; x.B_rim = x.b_rim * x.rho;
        }
        // which integrals this point needs
        int want = 0;
        if (valid) {
            const bool vel_on = !((x.N_ice < e) || (x.L_ice < e));                  // P3_terminal_velocity.jl:79-81
            const bool ice_on = (MODE == MODE_VEL) ? false : (x.q_ice > e && x.n_ice > e);   // BMT:961
            if (vel_on) want |= P3_WANT_VEL;
            if (ice_on) want |= P3_WANT_AGG | P3_WANT_COLL | ((x.T > a.tk.T_freeze) ? P3_WANT_MELT : 0);
            want &= a.want;
        }
        // The list is ordered by the LENGTH of a point's pass: an unrimed point has 2 non-empty mass-regime segments instead of 4
        // (half the outer quadrature nodes of every integral), a point without rain skips the closed-form rain integrals and
        // their root search.  An iteration lasts as long as its slowest warp (the barrier), so iterations made of short points
        // only are short; mixed in, the short points just wait.
        const int cat = ((x.L_rim > e && x.B_rim > e) ? 0 : 1) + ((x.L_rai > e && x.N_rai > e) ? 0 : 2);
        unsigned mbc[kCats];
#pragma unroll
        for (int c = 0; c < kCats; ++c) mbc[c] = __ballot_sync(0xffffffffu, want != 0 && cat == c);
        if (lane < kCats) warp_cnt[lane * W + warp] = __popc(mbc[lane]);
        __syncthreads();   // also: the previous round's slot reads are done
        int offset = 0, total = 0;
        for (int c = 0; c < kCats; ++c)
            for (int w = 0; w < W; ++w) {
                const int cnt = warp_cnt[c * W + w];
                offset += (c < cat || (c == cat && w < warp)) ? cnt : 0;
                total += cnt;
            }
        if (want) {
            idx[offset + __popc(mbc[cat] & ((1u << lane) - 1u))] = (unsigned short)threadIdx.x;
            wantv[threadIdx.x] = (unsigned char)want;
            double* sl = slots + threadIdx.x * kSlot;
            sl[0] = x.rho; sl[1] = x.T; sl[2] = x.L_ice; sl[3] = x.N_ice; sl[4] = x.L_rim; sl[5] = x.B_rim; sl[6] = x.logl;
            sl[7] = x.L_lcl; sl[8] = x.N_lcl; sl[9] = x.L_rai; sl[10] = x.N_rai;
        }
        // ---- the quantile solves of the round's listed points, ONE PER THREAD (a Halley iteration is serial: solved inside the
        // point's pass it kept 6 — velocities: 2 — lanes of the evaluating warp busy and the others waiting)
        __syncthreads();
        if (a.solve_logl) {   // §8(f)-1: the logλ solve fused into the P3 call (a 10-step Brent search over 72 gamma_inc per step)
            for (int j = threadIdx.x; j < total; j += BLOCK) {
                double* sl = slots + idx[j] * kSlot;
                sl[6] = p3_solve_logl(a.k, sl[2], sl[3], sl[4], sl[5], a.k.brent_iters);
            }
            __syncthreads();
        }
        constexpr int NQ = (MODE == MODE_VEL) ? 2 : 6;
        for (int task = threadIdx.x; task < total * NQ; task += BLOCK) {
            const int j = task / NQ, which = task - j * NQ;
            const int o = idx[j];
            double mu, lam;
            p3_mu_lam(a.k, slots[o * kSlot + 6], mu, lam);
            slots[o * kSlot + 11 + which] = p3_quantile(a.k, wantv[o], which, mu, lam);
        }
        for (int it = 0; it * W < total; ++it) {
            // it = 0: the list and the quantiles are complete; every CUMICRO_P3_SYNC_EVERY-th it: the block's warps start their next point together
            if (CUMICRO_P3_SYNC_EVERY == 1 || it % CUMICRO_P3_SYNC_EVERY == 0) __syncthreads();
            const int j = it * W + warp;
            if (j < total) {
                const int o = idx[j];
                double* sl = slots + o * kSlot;
                const double rho = sl[0], T = sl[1], L_ice = sl[2], N_ice = sl[3], L_rim = sl[4], B_rim = sl[5], logl = sl[6], L_lcl = sl[7],
                             N_lcl = sl[8], L_rai = sl[9], N_rai = sl[10];
                const int w_o = wantv[o];
                double pre = 0.0;
                if (lane < NQ) pre = sl[11 + lane];
                __syncwarp();
                P3Point& s = *reinterpret_cast<P3Point*>(points + warp * kPointDoubles);
                p3_point_init(s, a.p, a.k, rho, T, L_ice, N_ice, L_rim, B_rim, logl);   // every lane stores the same values
                __syncwarp();
                P3Rates r;
                p3_point_rates(s, a.p, a.k, a.tk, a.sk, qx, qw, sc, w_o, L_lcl, N_lcl, L_rai, N_rai, r, true, pre);
                if (lane == 0) {
                    sl[0] = r.v_n; sl[1] = r.v_m; sl[2] = r.melt_dN; sl[3] = r.melt_dL; sl[4] = r.agg_dN;
#pragma unroll
                    for (int c = 0; c < 7; ++c) sl[5 + c] = r.src[c];
                    sl[12] = s.F_rim; sl[13] = s.rho_rim;
                }
            } else {   // no point left for this warp: keep the block's barrier count (cm_p3.cuh, P3_BAR)
                for (int b = 0; b < p3_phase_barriers(nq); ++b) P3_BAR();
            }
        }
        __syncthreads();
        P3Rates mine;
        mine.v_n = mine.v_m = mine.melt_dN = mine.melt_dL = mine.agg_dN = 0.0;
#pragma unroll
        for (int c = 0; c < 7; ++c) mine.src[c] = 0.0;
        double F_rim = 0.0, rho_rim = 0.0;
        if (want) {
            const double* sl = slots + threadIdx.x * kSlot;
            mine.v_n = sl[0]; mine.v_m = sl[1]; mine.melt_dN = sl[2]; mine.melt_dL = sl[3]; mine.agg_dN = sl[4];
#pragma unroll
            for (int c = 0; c < 7; ++c) mine.src[c] = sl[5 + c];
            F_rim = sl[12]; rho_rim = sl[13];
        }
        if (!valid) continue;
        if (MODE == MODE_BMT) {
            if (want == 0) {   // state_from_prognostic for the rim bookkeeping of the pointwise processes (BMT:930)
                F_rim = fmin_(regularised_ratio_(fmin_(x.L_rim, x.L_ice), x.L_ice, e), 1.0 - e);
                rho_rim = fmin_(regularised_ratio_(x.L_rim, x.B_rim, e), a.k.rho_l08);
            }
            double y[9];
            bmt2m_p3_assemble(a.p, a.tk, a.sk, a.k, x, (x.q_ice > e && x.n_ice > e), mine, F_rim, rho_rim, y);
#pragma unroll
            for (int c = 0; c < 9; ++c)
                if (a.out[c]) a.out[c][i] = (FT)y[c];
        } else if (MODE == MODE_RATES) {
            const double y[12] = {mine.v_n, mine.v_m, mine.melt_dN, mine.melt_dL, mine.agg_dN, mine.src[0], mine.src[1],
                                  mine.src[2], mine.src[3], mine.src[4], mine.src[5], mine.src[6]};
#pragma unroll
            for (int c = 0; c < 12; ++c)
                if (a.out[c]) a.out[c][i] = (FT)y[c];
        } else {
            if (a.out[0]) a.out[0][i] = (FT)mine.v_n;
            if (a.out[1]) a.out[1][i] = (FT)mine.v_m;
        }
    }
#else
    const int64_t n_tiles = (a.n + 31) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * BLOCK) >> 5;
#if CUMICRO_P3_SYNC
    // block-uniform trip count: the point loop below holds a block barrier
    for (int64_t tile0 = ((int64_t)blockIdx.x * BLOCK) >> 5; tile0 < n_tiles; tile0 += n_warps) {
        const int64_t tile = tile0 + (threadIdx.x >> 5);
        const int64_t i = (tile << 5) + lane;
        const bool valid = tile < n_tiles && i < a.n;
#else
    for (int64_t tile = ((int64_t)blockIdx.x * BLOCK + threadIdx.x) >> 5; tile < n_tiles; tile += n_warps) {
        const int64_t i = (tile << 5) + lane;
        const bool valid = i < a.n;
#endif
        Pt x;
        auto ld = [&](int c) { return (valid && a.in[c]) ? (double)__ldg(a.in[c] + i) : 0.0; };
        if (MODE == MODE_VEL) {   // in: rho_a, L_ice, N_ice, L_rim, B_rim, logl (volumetric, as the reference's wrapper takes them)
            x.rho = ld(0); x.T = 273.15; x.q_tot = 0.0; x.q_lcl = x.n_lcl = x.q_rai = x.n_rai = 0.0;
            x.L_ice = ld(1); x.N_ice = ld(2); x.L_rim = ld(3); x.B_rim = ld(4); x.logl = ld(5); x.shift = 0.0;
            x.q_ice = x.n_ice = x.q_rim = x.b_rim = 0.0;
            x.L_lcl = x.N_lcl = x.L_rai = x.N_rai = 0.0;
        } else {                  // in: rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, q_rim, b_rim, logl [, inpc_log_shift]
            x.rho = clamp0_(ld(0)); x.T = ld(1); x.q_tot = clamp0_(ld(2)); x.q_lcl = clamp0_(ld(3)); x.n_lcl = clamp0_(ld(4));
            x.q_rai = clamp0_(ld(5)); x.n_rai = clamp0_(ld(6)); x.q_ice = clamp0_(ld(7)); x.n_ice = clamp0_(ld(8));
            x.q_rim = clamp0_(ld(9)); x.b_rim = clamp0_(ld(10)); x.logl = ld(11); x.shift = ld(12);
            x.L_lcl = x.q_lcl * x.rho; x.L_rai = x.q_rai * x.rho; x.N_lcl = x.n_lcl * x.rho; x.N_rai = x.n_rai * x.rho;
            x.L_ice = x.q_ice * x.rho; x.N_ice = x.n_ice * x.rho; x.L_rim = x.q_rim * x.rho; x.B_rim = x.b_rim * x.rho;
        }
        // which integrals this point needs
        int want = 0;
        if (valid) {
            const bool vel_on = !((x.N_ice < e) || (x.L_ice < e));                  // P3_terminal_velocity.jl:79-81
            const bool ice_on = (MODE == MODE_VEL) ? false : (x.q_ice > e && x.n_ice > e);   // BMT:961
            if (vel_on) want |= P3_WANT_VEL;
            if (ice_on) want |= P3_WANT_AGG | P3_WANT_COLL | ((x.T > a.tk.T_freeze) ? P3_WANT_MELT : 0);
            want &= a.want;
        }
        P3Rates mine;
        mine.v_n = mine.v_m = mine.melt_dN = mine.melt_dL = mine.agg_dN = 0.0;
#pragma unroll
        for (int c = 0; c < 7; ++c) mine.src[c] = 0.0;
        double F_rim = 0.0, rho_rim = 0.0;
        unsigned m = __ballot_sync(0xffffffffu, want != 0);
#if CUMICRO_P3_SYNC
        // every warp of the block starts its next point together: the SM's warps walk the same code at the same time
        while (__syncthreads_or(m != 0)) {
            if (m == 0) {   // no point left in this warp's tile: keep the block's barrier count (cm_p3.cuh, P3_BAR)
                for (int b = 0; b < p3_phase_barriers(nq); ++b) P3_BAR();
                continue;
            }
#else
        while (m) {
#endif
            const int b = __ffs(m) - 1;
            m &= m - 1;
            P3Point s;
            p3_point_init(s, a.p, a.k, bcast(x.rho, b), bcast(x.T, b), bcast(x.L_ice, b), bcast(x.N_ice, b), bcast(x.L_rim, b),
                          bcast(x.B_rim, b), bcast(x.logl, b));
            P3Rates r;
            p3_point_rates(s, a.p, a.k, a.tk, a.sk, qx, qw, sc, __shfl_sync(0xffffffffu, want, b), bcast(x.L_lcl, b), bcast(x.N_lcl, b),
                           bcast(x.L_rai, b), bcast(x.N_rai, b), r);
            if (lane == b) { mine = r; F_rim = s.F_rim; rho_rim = s.rho_rim; }
        }
        if (!valid) continue;
        if (MODE == MODE_BMT) {
            if (want == 0) {   // state_from_prognostic for the rim bookkeeping of the pointwise processes (BMT:930)
                F_rim = fmin_(regularised_ratio_(fmin_(x.L_rim, x.L_ice), x.L_ice, e), 1.0 - e);
                rho_rim = fmin_(regularised_ratio_(x.L_rim, x.B_rim, e), a.k.rho_l08);
            }
            double y[9];
            bmt2m_p3_assemble(a.p, a.tk, a.sk, a.k, x, (x.q_ice > e && x.n_ice > e), mine, F_rim, rho_rim, y);
#pragma unroll
            for (int c = 0; c < 9; ++c)
                if (a.out[c]) a.out[c][i] = (FT)y[c];
        } else if (MODE == MODE_RATES) {
            const double y[12] = {mine.v_n, mine.v_m, mine.melt_dN, mine.melt_dL, mine.agg_dN, mine.src[0], mine.src[1],
                                  mine.src[2], mine.src[3], mine.src[4], mine.src[5], mine.src[6]};
#pragma unroll
            for (int c = 0; c < 12; ++c)
                if (a.out[c]) a.out[c][i] = (FT)y[c];
        } else {
            if (a.out[0]) a.out[0][i] = (FT)mine.v_n;
            if (a.out[1]) a.out[1][i] = (FT)mine.v_m;
        }
    }
#endif
}

template <class FT> struct PP3;
template <> struct PP3<double> { using type = cumicro_params_p3_f64; };
template <> struct PP3<float> { using type = cumicro_params_p3_f32; };
template <class FT> constexpr bool is_f32() { return sizeof(FT) == 4; }

template <class FT> int p3_check(const typename PP3<FT>::type* p) {
    if (p == nullptr) return cmh::fail(CUMICRO_E_NULL, "parameter block is NULL");
    if (p->quad.n < 1 || p->quad.n > kQuadMax) return cmh::fail(CUMICRO_E_OPTION, "quad.n = %d (expected 1..%d)", (int)p->quad.n, kQuadMax);
    if (p->warm.sb.pdf_r.limited != 0 && p->warm.sb.pdf_r.limited != 1) return cmh::fail(CUMICRO_E_OPTION, "sb.pdf_r.limited = %d", (int)p->warm.sb.pdf_r.limited);
    if (p->scheme.slope_power_law != 0 && p->scheme.slope_power_law != 1) return cmh::fail(CUMICRO_E_OPTION, "scheme.slope_power_law = %d", (int)p->scheme.slope_power_law);
    if (p->scheme.aspect_oblate != 0 && p->scheme.aspect_oblate != 1) return cmh::fail(CUMICRO_E_OPTION, "scheme.aspect_oblate = %d", (int)p->scheme.aspect_oblate);
    // P3_processes.jl:616: @assert ρw == psd_r.ρw
    if (p->warm.sb.pdf_c.rho_w != p->warm.sb.pdf_r.rho_w) return cmh::fail(CUMICRO_E_OPTION, "cloud and rain PSDs must share the liquid water density");
    return CUMICRO_OK;
}

template <class FT, int MODE>
int p3_launch(const typename PP3<FT>::type* p, int64_t n, const FT* const* in, int nin, int nin_required, FT* const* out, int nout, int want,
              void* stream, const char* what) {
    int st = p3_check<FT>(p);
    if (st) return st;
    if (n < 0) return cmh::fail(CUMICRO_E_SIZE, "n = %lld is negative", (long long)n);
    if (!in || !out) return cmh::fail(CUMICRO_E_NULL, "%s: column pointer table is NULL", what);
    const int logl_col = (MODE == MODE_VEL) ? 5 : 11;   // may be NULL: the kernel then solves logλ itself
    for (int c = 0; c < nin_required; ++c)
        if (n > 0 && in[c] == nullptr && c != logl_col) return cmh::fail(CUMICRO_E_NULL, "%s: input column %d is NULL", what, c);
    if (n == 0) return CUMICRO_OK;
    P3Args<FT> a{};
    widen(*p, a.p);
    a.tk = make_thermo_k<double>(a.p.warm.tps, is_f32<FT>());
    a.sk = make_sb2006_k<double>(a.p.warm.sb, a.p.warm.aps, is_f32<FT>());
    a.k = make_p3_k(a.p, is_f32<FT>());
    for (int c = 0; c < NIN_MAX; ++c) a.in[c] = (c < nin) ? in[c] : nullptr;
    for (int c = 0; c < NOUT_MAX; ++c) a.out[c] = (c < nout) ? out[c] : nullptr;
    a.n = n;
    a.want = want;
    a.solve_logl = (in[logl_col] == nullptr) ? 1 : 0;
    const int nq = a.k.n;
    size_t shmem = sizeof(double) * (size_t)(2 * nq + (BLOCK / 32) * P3Scratch::doubles(nq));
#if CUMICRO_P3_SYNC
    shmem += sizeof(double) * BLOCK * kSlot + BLOCK * 3 + 16 + (BLOCK / 32) * ((sizeof(P3Point) + 7) / 8) * sizeof(double);
#endif
    const int64_t tiles = (n + 31) / 32;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((tiles + BLOCK / 32 - 1) / (BLOCK / 32), (int64_t)cmh::num_sms() * MINB));
    auto kern = p3_tile_kernel<FT, MODE>;
    if (shmem + 8 * 1024 > 48 * 1024) {   // dynamic + the static math tables (4 KB) against the 48 KB default
        st = cmh::cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem), "cudaFuncSetAttribute");
        if (st) return st;
    }
    kern<<<blocks, BLOCK, shmem, (cudaStream_t)stream>>>(a);
    cmh::count_launch();
    return cmh::cuda_status(cudaGetLastError(), what);
}

// ---- the F23 / Bigg nucleation rates on their own: one point per thread --------------------------------------
template <bool WITH_SHIFT> struct F23Functor {
    cumicro_params_p3_f64 p;
    ThermoK<double> tk;
    SB2006K<double> sk;
    P3K k;
    __device__ __forceinline__ void operator()(const double (&v)[WITH_SHIFT ? 10 : 9], double (&y)[7]) const {
        Pt x{};
        x.rho = clamp0_(v[0]); x.T = v[1]; x.q_tot = clamp0_(v[2]); x.q_lcl = clamp0_(v[3]); x.n_lcl = clamp0_(v[4]);
        x.q_rai = clamp0_(v[5]); x.n_rai = clamp0_(v[6]); x.q_ice = clamp0_(v[7]); x.n_ice = clamp0_(v[8]); x.shift = WITH_SHIFT ? v[WITH_SHIFT ? 9 : 0] : 0.0;
        x.N_lcl = x.n_lcl * x.rho; x.N_rai = x.n_rai * x.rho;
        const F23Rates f = f23_rates(p, tk, sk, k, x);
        y[0] = f.rain_dn; y[1] = f.rain_dq; y[2] = f.cld_dn; y[3] = f.cld_dq; y[4] = f.cap_dn; y[5] = f.dep_dn; y[6] = f.dep_dq;
    }
};

template <class FT>
int icenuc_f23_impl(const typename PP3<FT>::type* p, int64_t n, const FT* const* in9, const FT* shift, FT* const* out7, void* stream) {
    int st = p3_check<FT>(p);
    if (st) return st;
    if (!in9 || !out7) return cmh::fail(CUMICRO_E_NULL, "icenuc_f23: column pointer table is NULL");
    if (n < 0) return cmh::fail(CUMICRO_E_SIZE, "n = %lld is negative", (long long)n);
    if (n == 0) return CUMICRO_OK;
    const FT* in[10];
    for (int c = 0; c < 9; ++c) {
        if (in9[c] == nullptr) return cmh::fail(CUMICRO_E_NULL, "icenuc_f23: input column %d is NULL", c);
        in[c] = in9[c];
    }
    FT* out[7];
    for (int c = 0; c < 7; ++c) out[c] = out7[c];
    cudaStream_t s = (cudaStream_t)stream;
    auto fill = [&](auto& f) {
        widen(*p, f.p);
        f.tk = make_thermo_k<double>(f.p.warm.tps, is_f32<FT>());
        f.sk = make_sb2006_k<double>(f.p.warm.sb, f.p.warm.aps, is_f32<FT>());
        f.k = make_p3_k(f.p, is_f32<FT>());
    };
    if (shift) {
        in[9] = shift;
        F23Functor<true> f{};
        fill(f);
        return launch_pointwise_tiled<FT, 10, 7, F23Functor<true>, 128, 5>(f, n, in, out, s, "icenuc_f23 launch");
    }
    const FT* in9c[9];
    for (int c = 0; c < 9; ++c) in9c[c] = in[c];
    F23Functor<false> f{};
    fill(f);
    return launch_pointwise_tiled<FT, 9, 7, F23Functor<false>, 128, 5>(f, n, in9c, out, s, "icenuc_f23 launch");
}

// ---- P3.get_distribution_logλ_from_prognostic: one point per thread ------------------------------
// P3.state_from_prognostic + P3State thresholds                     P3_particle_properties.jl:43-56, 101-106, 191-272
struct P3Thresholds { double F_rim, rho_rim, rho_g, D_gr, D_cr; };
__device__ inline P3Thresholds p3_thresholds(const P3K& k, double L_ice, double L_rim, double B_rim) {
    P3Thresholds t;
    t.F_rim = fmin_(regularised_ratio_(fmin_(L_rim, L_ice), L_ice, k.eps), 1.0 - k.eps);
    t.rho_rim = fmin_(regularised_ratio_(L_rim, B_rim, k.eps), k.rho_l08);
    const double pp = k.thr_p;
    const double logFu = log1p_(-t.F_rim);
    auto exprel1 = [](double v) { return expm1_(v) / v; };
    auto exprel2 = [](double v) {
        if (fabs(v) < 0.2) {
            double r = 1.0 / 362880.0;
            const double c[7] = {1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 1.0 / 2.0};
#pragma unroll
            for (int i = 0; i < 7; ++i) r = r * v + c[i];
            return r;
        }
        return (expm1_(v) - v) / (v * v);
    };
    const double phi1 = exprel1(logFu), phi1mp = exprel1((1.0 - pp) * logFu);
    const double H = -pp * exprel2(-pp * logFu) - (1.0 - pp) * exprel2((1.0 - pp) * logFu);
    const double rho_d = -(t.rho_rim * phi1 * phi1mp) / (H - phi1mp * phi1);
    t.rho_g = t.F_rim * t.rho_rim + (1.0 - t.F_rim) * rho_d;
    const bool unrimed = (t.F_rim == 0.0);
    const double pi = num<double>::pi();
    t.D_gr = unrimed ? num<double>::inf() : pow_pos_(k.thr_coef / (pi * t.rho_g), pp);
    t.D_cr = unrimed ? num<double>::inf() : pow_pos_(k.thr_coef / (pi * (t.rho_g * (1.0 - t.F_rim))), pp);
    return t;
}
// get_μ and logmass_gamma_moment(state, μ, logλ; n)                   P3_size_distribution.jl:171, 193-200
__device__ inline double p3_logmass_moment(const P3K& k, const P3Thresholds& t, double logl, double n_mom, double& mu) {
    const double lam = exp_full_(logl);
    mu = k.slope_power_law ? clamp_(k.slope_a * pow_pos_(lam, k.slope_b) - k.slope_c, 0.0, k.mu_max) : k.mu_const;
    const double pi = num<double>::pi();
    const double inf = num<double>::inf();
    const double bnd[5] = {0.0, clamp_(k.D_th, 0.0, inf), clamp_(t.D_gr, 0.0, inf), clamp_(t.D_cr, 0.0, inf), inf};
    const double Fu = fmax_(1.0 - t.F_rim, k.eps);
    double m[4];
#pragma unroll
    for (int sgm = 0; sgm < 4; ++sgm) {
        const double D1 = bnd[sgm], D2 = bnd[sgm + 1];
        const double Dm = (D1 + D2) / 2.0;
        const int r = (Dm < k.D_th) ? 0 : ((t.F_rim == 0.0) ? 1 : ((Dm < t.D_gr) ? 2 : ((Dm < t.D_cr) ? 3 : 4)));
        const double a = (r == 0) ? k.rho_i * pi / 6.0 : ((r == 3) ? t.rho_g * pi / 6.0 : ((r == 4) ? k.alpha_va / Fu : k.alpha_va));
        const double b = (r == 0 || r == 3) ? 3.0 : k.beta_va;
        if (!(D1 < D2)) { m[sgm] = -inf; continue; }
        const double z = (b + n_mom) + mu + 1.0;
        const double x1 = D1 * lam, x2 = D2 * lam;
        const double lg = lgamma_pos_(z);
        const PQ g1 = gamma_inc_(z, x1, lg, k.gamma_iters), g2 = gamma_inc_(z, x2, lg, k.gamma_iters);
        double dq = (x2 < z + 1.0) ? g2.P - g1.P : g1.Q - g2.Q;
        dq = fmax_(dq, k.eps);
        m[sgm] = -z * logl + lg + log_full_(dq) + log_full_(a);
    }
    double mx = m[0];
#pragma unroll
    for (int i = 1; i < 4; ++i) mx = fmax_(mx, m[i]);
    if (!isfinite(mx)) return mx;
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) sum += exp_full_(m[i] - mx);
    return mx + log_full_(sum);
}

struct P3LogLambda {
    P3K k;
    int iters;   // Brent iterations (reference: 10 / 8)
    __device__ __forceinline__ void operator()(const double (&x)[4], double (&y)[1]) const { y[0] = p3_solve_logl(k, x[0], x[1], x[2], x[3], iters); }
};
__device__ __noinline__ double p3_solve_logl(const P3K& k, double L_ice, double N_ice, double L_rim, double B_rim, int iters) {
    const P3Thresholds t = p3_thresholds(k, L_ice, L_rim, B_rim);
    if (N_ice < k.eps || L_ice < k.eps) return -num<double>::inf();
    const double target = log_full_(L_ice) - log_full_(N_ice);
    // shape_problem(logλ) = logLdivN(state, logλ) - target                        P3_size_distribution.jl:211-216, 292
    auto f = [&](double l) {
        double mu;
        const double lse = p3_logmass_moment(k, t, l, 0.0, mu);
        const double z0 = 0.0 + mu + 1.0;
        return (lse - (-z0 * l + lgamma_pos_(z0) + 0.0)) - target;
    };
    const double lo = 2.0, hi = 17.0;
    const double f_lo = f(lo), f_hi = f(hi);
    if (!isfinite(f_lo) || !isfinite(f_hi) || f_lo * f_hi > 0.0) return (fabs(f_lo) <= fabs(f_hi)) ? lo : hi;
    return brent_fixed(f, lo, hi, f_lo, f_hi, iters);
}

// P3State thresholds and the mass-weighted mean diameter D_m of (state, logλ)     P3_integral_properties.jl:56-61
struct P3StateDiag {
    P3K k;
    __device__ __forceinline__ void operator()(const double (&x)[5], double (&y)[7]) const {
        const double L_ice = x[0], N_ice = x[1], L_rim = x[2], B_rim = x[3], logl = x[4];
        const P3Thresholds t = p3_thresholds(k, L_ice, L_rim, B_rim);
        y[0] = t.F_rim; y[1] = t.rho_rim; y[2] = t.rho_g; y[3] = k.D_th; y[4] = t.D_gr; y[5] = t.D_cr;
        double mu;
        const double lse = p3_logmass_moment(k, t, logl, 1.0, mu);
        const double z0 = 0.0 + mu + 1.0;
        const double logN0 = log_full_(N_ice) - (-z0 * logl + lgamma_pos_(z0) + 0.0);
        y[6] = exp_full_(logN0 + lse) / L_ice;
    }
};

// UT.gamma_inc / gamma_inc_inv / rime_mass_fraction / rime_density over columns (the reference tests them on the
// device, test/gpu_tests.jl:1305-1338)
struct P3Leaf {
    int what, gamma_iters;
    double eps;
    __device__ __forceinline__ void operator()(const double (&x)[2], double (&y)[1]) const {
        switch (what) {
            case 0: y[0] = gamma_inc_(x[0], x[1], lgamma_pos_(x[0]), gamma_iters).P; break;
            case 1: y[0] = gamma_inc_(x[0], x[1], lgamma_pos_(x[0]), gamma_iters).Q; break;
            case 2: y[0] = gamma_inc_inv_(x[0], x[1], 1.0 - x[1], gamma_iters, eps); break;
            case 3: y[0] = regularised_ratio_(fmin_(x[0], x[1]), x[1], eps); break;
            case 4: y[0] = regularised_ratio_(x[0], x[1], eps); break;
            default: y[0] = 0.0;
        }
    }
};

template <class FT>
int p3_logl_impl(const typename PP3<FT>::type* p, int64_t n, const FT* L_ice, const FT* N_ice, const FT* L_rim, const FT* B_rim, int iters,
                 FT* logl, void* stream) {
    int st = p3_check<FT>(p);
    if (st) return st;
    const FT* in[4] = {L_ice, N_ice, L_rim, B_rim};
    if ((st = validate_columns<FT, 4>(p, n, in))) return st;
    FT* out[1] = {logl};
    if ((st = require_outputs<FT, 1>(n, out, 1))) return st;
    cumicro_params_p3_f64 wide;
    widen(*p, wide);
    P3LogLambda f{};
    f.k = make_p3_k(wide, is_f32<FT>());
    f.iters = iters > 0 ? iters : f.k.brent_iters;
    return launch_pointwise<FT, 4, 1, P3LogLambda, CUMICRO_P3L_BLOCK, CUMICRO_P3L_MINB, false>(f, n, in, out, (cudaStream_t)stream, "p3_logl kernel launch");
}

template <class FT>
int p3_state_impl(const typename PP3<FT>::type* p, int64_t n, const FT* L_ice, const FT* N_ice, const FT* L_rim, const FT* B_rim,
                  const FT* logl, FT* const* out7, void* stream) {
    int st = p3_check<FT>(p);
    if (st) return st;
    const FT* in[5] = {L_ice, N_ice, L_rim, B_rim, logl};
    if ((st = validate_columns<FT, 5>(p, n, in))) return st;
    if (out7 == nullptr) return cmh::fail(CUMICRO_E_NULL, "p3_state: output pointer table is NULL");
    FT* out[7];
    for (int c = 0; c < 7; ++c) out[c] = out7[c];
    cumicro_params_p3_f64 wide;
    widen(*p, wide);
    P3StateDiag f{};
    f.k = make_p3_k(wide, is_f32<FT>());
    return launch_pointwise<FT, 5, 7, P3StateDiag, 128, 3, false>(f, n, in, out, (cudaStream_t)stream, "p3_state kernel launch");
}

template <class FT> int p3_leaf_impl(int what, int64_t n, const FT* x, const FT* y, FT* out, void* stream) {
    if (what < 0 || what > 4) return cmh::fail(CUMICRO_E_OPTION, "p3_leaf: what = %d (expected 0..4)", what);
    const FT* in[2] = {x, y};
    int st = validate_columns<FT, 2>(&what, n, in);
    if (st) return st;
    FT* o[1] = {out};
    if ((st = require_outputs<FT, 1>(n, o, 1))) return st;
    P3Leaf f{what, is_f32<FT>() ? 20 : 30, is_f32<FT>() ? 1.1920928955078125e-07 : 2.220446049250313e-16};
    return launch_pointwise<FT, 2, 1, P3Leaf, 128, 4, false>(f, n, in, o, (cudaStream_t)stream, "p3_leaf kernel launch");
}

}  // namespace

extern "C" {

int cumicro_p3_rates_f64(const cumicro_params_p3_f64* p, int64_t n, const double* const* in12, double* const* out12, void* stream) {
    return p3_launch<double, MODE_RATES>(p, n, in12, 12, 12, out12, 12, P3_WANT_VEL | P3_WANT_MELT | P3_WANT_AGG | P3_WANT_COLL, stream, "p3_rates");
}
int cumicro_p3_rates_f32(const cumicro_params_p3_f32* p, int64_t n, const float* const* in12, float* const* out12, void* stream) {
    return p3_launch<float, MODE_RATES>(p, n, in12, 12, 12, out12, 12, P3_WANT_VEL | P3_WANT_MELT | P3_WANT_AGG | P3_WANT_COLL, stream, "p3_rates");
}
int cumicro_bmt2m_p3_f64(const cumicro_params_p3_f64* p, int64_t n, const double* const* in12, const double* inpc_log_shift,
                         double* const* out9, void* stream) {
    if (!in12) return cmh::fail(CUMICRO_E_NULL, "bmt2m_p3: column pointer table is NULL");
    const double* in[13];
    for (int c = 0; c < 12; ++c) in[c] = in12[c];
    in[12] = inpc_log_shift;
    return p3_launch<double, MODE_BMT>(p, n, in, 13, 12, out9, 9, P3_WANT_MELT | P3_WANT_AGG | P3_WANT_COLL, stream, "bmt2m_p3");
}
int cumicro_bmt2m_p3_f32(const cumicro_params_p3_f32* p, int64_t n, const float* const* in12, const float* inpc_log_shift,
                         float* const* out9, void* stream) {
    if (!in12) return cmh::fail(CUMICRO_E_NULL, "bmt2m_p3: column pointer table is NULL");
    const float* in[13];
    for (int c = 0; c < 12; ++c) in[c] = in12[c];
    in[12] = inpc_log_shift;
    return p3_launch<float, MODE_BMT>(p, n, in, 13, 12, out9, 9, P3_WANT_MELT | P3_WANT_AGG | P3_WANT_COLL, stream, "bmt2m_p3");
}
int cumicro_termvel_p3_f64(const cumicro_params_p3_f64* p, int64_t n, const double* rho_a, const double* L_ice, const double* N_ice,
                           const double* L_rim, const double* B_rim, const double* logl, double* v_n, double* v_m, void* stream) {
    const double* in[6] = {rho_a, L_ice, N_ice, L_rim, B_rim, logl};
    double* out[2] = {v_n, v_m};
    return p3_launch<double, MODE_VEL>(p, n, in, 6, 6, out, 2, P3_WANT_VEL, stream, "termvel_p3");
}
int cumicro_termvel_p3_f32(const cumicro_params_p3_f32* p, int64_t n, const float* rho_a, const float* L_ice, const float* N_ice,
                           const float* L_rim, const float* B_rim, const float* logl, float* v_n, float* v_m, void* stream) {
    const float* in[6] = {rho_a, L_ice, N_ice, L_rim, B_rim, logl};
    float* out[2] = {v_n, v_m};
    return p3_launch<float, MODE_VEL>(p, n, in, 6, 6, out, 2, P3_WANT_VEL, stream, "termvel_p3");
}
int cumicro_p3_logl_f64(const cumicro_params_p3_f64* p, int64_t n, const double* L_ice, const double* N_ice, const double* L_rim,
                        const double* B_rim, int brent_iters, double* logl, void* stream) {
    return p3_logl_impl<double>(p, n, L_ice, N_ice, L_rim, B_rim, brent_iters, logl, stream);
}
int cumicro_p3_logl_f32(const cumicro_params_p3_f32* p, int64_t n, const float* L_ice, const float* N_ice, const float* L_rim,
                        const float* B_rim, int brent_iters, float* logl, void* stream) {
    return p3_logl_impl<float>(p, n, L_ice, N_ice, L_rim, B_rim, brent_iters, logl, stream);
}

int cumicro_icenuc_f23_f64(const cumicro_params_p3_f64* p, int64_t n, const double* const* in9, const double* inpc_log_shift,
                           double* const* out7, void* stream) {
    return icenuc_f23_impl<double>(p, n, in9, inpc_log_shift, out7, stream);
}
int cumicro_icenuc_f23_f32(const cumicro_params_p3_f32* p, int64_t n, const float* const* in9, const float* inpc_log_shift,
                           float* const* out7, void* stream) {
    return icenuc_f23_impl<float>(p, n, in9, inpc_log_shift, out7, stream);
}

int cumicro_p3_state_f64(const cumicro_params_p3_f64* p, int64_t n, const double* L_ice, const double* N_ice, const double* L_rim,
                         const double* B_rim, const double* logl, double* const* out7, void* stream) {
    return p3_state_impl<double>(p, n, L_ice, N_ice, L_rim, B_rim, logl, out7, stream);
}
int cumicro_p3_state_f32(const cumicro_params_p3_f32* p, int64_t n, const float* L_ice, const float* N_ice, const float* L_rim,
                         const float* B_rim, const float* logl, float* const* out7, void* stream) {
    return p3_state_impl<float>(p, n, L_ice, N_ice, L_rim, B_rim, logl, out7, stream);
}
int cumicro_p3_leaf_f64(int what, int64_t n, const double* x, const double* y, double* out, void* stream) {
    return p3_leaf_impl<double>(what, n, x, y, out, stream);
}
int cumicro_p3_leaf_f32(int what, int64_t n, const float* x, const float* y, float* out, void* stream) {
    return p3_leaf_impl<float>(what, n, x, y, out, stream);
}

}  // extern "C"
