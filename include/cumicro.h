/* cumicro.h — C ABI of libcumicro.so: B200 (sm_100a) bulk cloud-microphysics
 * tendencies.
 *
 * This is the drop-in boundary for the data-parallel hot path of
 * CliMA/CloudMicrophysics.jl (reference file:line cited per entry point).  The
 * reference exposes pointwise Julia methods that a host model broadcasts over
 * arrays; each entry point below is the array-level form of one such method:
 * structure-of-arrays columns in, structure-of-arrays columns out, one fused
 * kernel per call.  A Julia package extension packs the
 * live `mp`/`tps` parameter objects into the POD blocks declared in
 * cumicro_params.inc and `ccall`s these symbols.
 *
 * Conventions
 *  - Plain C types only.  `_f64` / `_f32` suffix = Float64 / Float32 method.
 *  - Device entry points: every column pointer is a DEVICE pointer to `n`
 *    contiguous elements, owned by the caller; work is enqueued on `stream`
 *    (a cudaStream_t passed as void*, NULL = legacy default stream) and the call
 *    returns without synchronising.  The caller selects the device.
 *  - `_host` entry points take HOST pointers (pinned memory recommended), run
 *    a chunked H2D -> kernel -> D2H pipeline on internal streams and return
 *    after the results are in the host buffers.
 *  - Output pointers documented as "optional" may be NULL (column skipped).
 *  - Return value: 0 ok; <0 API misuse (CUMICRO_E_*); >0 a cudaError_t value.
 *    cumicro_last_error() returns a thread-local message for the last failure.
 *  - Input domain: finite columns with rho > 0 and T > 0; contents / numbers of any sign (clamped like the reference), zeros and
 *    sub-eps values anywhere.  A non-finite value or rho = 0 makes THAT cell's outputs unspecified ("Input domain").
 *  - Per-point domain violations (the reference throws DomainError /
 *    AssertionError, IN:558-562, IN:47,73) cannot throw from a kernel: the
 *    output is NaN and a device counter is incremented (see *_status args).
 *  - The library never falls back to the CPU.
 */
#ifndef CUMICRO_H
#define CUMICRO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUMICRO_VERSION 100 /* 0.1.0 */

#define CUMICRO_OK 0
#define CUMICRO_E_NULL (-1)      /* required pointer is NULL */
#define CUMICRO_E_SIZE (-2)      /* n < 0 or size overflow */
#define CUMICRO_E_OPTION (-3)    /* unknown option / enum value in a parameter block */
#define CUMICRO_E_NODEVICE (-4)  /* no CUDA device / driver */
#define CUMICRO_E_ARG (-5)       /* other invalid argument */

/* Option values of cumicro_options_1m_* (one slot per process; see cumicro_params.inc). */
enum {
    CUMICRO_1M_OFF = 0,
    CUMICRO_1M_CLOUD_ICE_CONSTANT_TIMESCALE = 1, /* ConstantTimescale */
    CUMICRO_1M_CLOUD_ICE_TEMPERATURE_DEPENDENT = 2, /* TemperatureDependent */
    CUMICRO_1M_RAIN_ACNV_KESSLER = 1, /* Kessler1M */
    CUMICRO_1M_RAIN_ACNV_PRESCRIBED_ND = 2, /* PrescribedNd */
    CUMICRO_1M_SNOW_ACNV_NO_SUPERSAT = 1, /* NoSupersaturation */
    CUMICRO_1M_SNOW_ACNV_WITH_SUPERSAT = 2, /* WithSupersaturation */
    CUMICRO_1M_SNOW_SUBLIMATION_ONLY = 1, /* SublimationOnly */
    CUMICRO_1M_SNOW_DEPOSITION_AND_SUBLIMATION = 2, /* DepositionAndSublimation */
    CUMICRO_1M_ON = 1
};

#define CUMICRO_FT double
#define CUMICRO_T(x) x##_f64
#include "cumicro_params.inc"
#undef CUMICRO_FT
#undef CUMICRO_T

#define CUMICRO_FT float
#define CUMICRO_T(x) x##_f32
#include "cumicro_params.inc"
#undef CUMICRO_FT
#undef CUMICRO_T

int cumicro_version(void);
const char* cumicro_last_error(void);

/* Number of kernels launched by this library in the calling process so far
 * (monotonic; used by bench.py for `gpu_launches`). */
int64_t cumicro_launch_count(void);

/* Roofline probe: enqueue `iters` x 8 independent FP64 FMAs per thread on
 * SMs x blocks_per_sm blocks of 256 threads; *flops_out = flops issued (2 per FMA).
 * bench.py times it with CUDA events to obtain the FP64 pipe peak in place. */
int cumicro_probe_fp64_fma(int64_t iters, int blocks_per_sm, double* scratch, double* flops_out, void* stream);

/* Accuracy probe of the library's device math (cm_math.cuh) on the GPU itself:
 * fn = 0 exp, 1 log, 2 cbrt, 3 reciprocal, 4 pow(x, y), 5 exp with IEEE limits, 6 sqrt, 7 erf, 8 log1p (x > 0),
 *      9 x^(-1/3), 10 x / y through the shared correctly rounded reciprocal;
 * out[i] = fn(x[i] [, y[i]]).  Device pointers; used by tests/test_gpu_math.py. */
int cumicro_probe_math_f64(int fn, int64_t n, const double* x, const double* y, double* out, void* stream);

/* Free the per-thread device staging buffers that the `_host` entry points cache. */
void cumicro_release_workspace(void);

/* ---------------------------------------------------------------------------
 * 2-moment warm rain (Seifert-Beheng 2006)
 * Replaces: BulkMicrophysicsTendencies.bulk_microphysics_tendencies(
 *     ::Microphysics2Moment, mp::Microphysics2MParams{WR,Nothing}, tps,
 *     rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai)       BMT:820-854
 * (body: warm_rain_tendencies_2m, BMT:707-782).
 * `q_ice` is the optional 8th positional argument of that method (BMT:823): the
 * cloud-ice content seen by the thermodynamics; NULL means zero.
 * Outputs: dq_lcl_dt, dn_lcl_dt, dq_rai_dt, dn_rai_dt (required);
 * zero4[0..3] = dq_ice_dt, dq_rim_dt, db_rim_dt, dn_lcl_activation_dt are
 * identically zero in the reference (BMT:840-853): optional, filled with 0.
 * ------------------------------------------------------------------------- */
int cumicro_bmt2m_warm_f64(const cumicro_params_2m_warm_f64* p, int64_t n,
                           const double* rho, const double* T, const double* q_tot,
                           const double* q_lcl, const double* n_lcl,
                           const double* q_rai, const double* n_rai,
                           const double* q_ice /* optional: NULL = 0 */,
                           double* dq_lcl_dt, double* dn_lcl_dt,
                           double* dq_rai_dt, double* dn_rai_dt,
                           double* const* zero4, void* stream);
int cumicro_bmt2m_warm_f32(const cumicro_params_2m_warm_f32* p, int64_t n,
                           const float* rho, const float* T, const float* q_tot,
                           const float* q_lcl, const float* n_lcl,
                           const float* q_rai, const float* n_rai,
                           const float* q_ice /* optional: NULL = 0 */,
                           float* dq_lcl_dt, float* dn_lcl_dt,
                           float* dq_rai_dt, float* dn_rai_dt,
                           float* const* zero4, void* stream);

/* Same tendencies through HOST buffers (e2e path).  `chunk` = points per
 * pipeline stage (0 = library default). */
int cumicro_bmt2m_warm_host_f64(const cumicro_params_2m_warm_f64* p, int64_t n,
                                const double* rho, const double* T, const double* q_tot,
                                const double* q_lcl, const double* n_lcl,
                                const double* q_rai, const double* n_rai,
                                double* dq_lcl_dt, double* dn_lcl_dt,
                                double* dq_rai_dt, double* dn_rai_dt, int64_t chunk);
int cumicro_bmt2m_warm_host_f32(const cumicro_params_2m_warm_f32* p, int64_t n,
                                const float* rho, const float* T, const float* q_tot,
                                const float* q_lcl, const float* n_lcl,
                                const float* q_rai, const float* n_rai,
                                float* dq_lcl_dt, float* dn_lcl_dt,
                                float* dq_rai_dt, float* dn_rai_dt, int64_t chunk);

/* Individual SB2006 process rates, one column each (the array form of the
 * leaf methods the reference exports and tests one by one,
 * test/gpu_tests.jl:220-235).  `out` is a HOST array of CUMICRO_SB2006_NLEAF
 * device column pointers (NULL entries skipped), in this order: */
enum {
    CUMICRO_SB_COND_DQ_LCL = 0, /* NEQ._conv_q_vap_to_q_lcl_const   NEQ:117-140 */
    CUMICRO_SB_EVAP_DN_RAI,     /* CM2.rain_evaporation d(rho n_rai)/dt CM2:780-828 */
    CUMICRO_SB_EVAP_DQ_RAI,     /* CM2.rain_evaporation dq_rai/dt */
    CUMICRO_SB_ACNV_DQ_LCL,     /* CM2.autoconversion  CM2:396-427 */
    CUMICRO_SB_ACNV_DN_LCL,     /*   dN_lcl_dt [1/m3/s] */
    CUMICRO_SB_ACNV_DQ_RAI,
    CUMICRO_SB_ACNV_DN_RAI,
    CUMICRO_SB_LCL_SELFCOL,     /* CM2.cloud_liquid_self_collection CM2:488-501 */
    CUMICRO_SB_ACCR_DQ_LCL,     /* CM2.accretion CM2:445-470 */
    CUMICRO_SB_ACCR_DN_LCL,
    CUMICRO_SB_ACCR_DQ_RAI,
    CUMICRO_SB_RAI_SELFCOL,     /* CM2.rain_self_collection CM2:545-560 */
    CUMICRO_SB_RAI_BREAKUP,     /* CM2.rain_breakup CM2:579-601 */
    CUMICRO_SB_NUMADJ_LCL,      /* CM2.number_tendency_from_mass_limits CM2:882-891 */
    CUMICRO_SB_NUMADJ_RAI,
    CUMICRO_SB2006_NLEAF
};
int cumicro_sb2006_leaves_f64(const cumicro_params_2m_warm_f64* p, int64_t n,
                              const double* rho, const double* T, const double* q_tot,
                              const double* q_lcl, const double* n_lcl,
                              const double* q_rai, const double* n_rai,
                              double* const* out, void* stream);
int cumicro_sb2006_leaves_f32(const cumicro_params_2m_warm_f32* p, int64_t n,
                              const float* rho, const float* T, const float* q_tot,
                              const float* q_lcl, const float* n_lcl,
                              const float* q_rai, const float* n_rai,
                              float* const* out, void* stream);

/* CM2.rain_evaporation(sb, aps, tps, q_tot, q_lcl, q_icl, q_rai, q_sno, rho, N_rai, T) on its own (CM2:780-828; the leaf takes its
 * arguments unclamped and N_rai as a number density [1/m3]) together with its leading-order derivatives
 * CM2.∂rain_evaporation_∂N_rai_∂q_rai (CM2:844-853).
 *   in8  = HOST array of 8 device columns in the reference's argument order: q_tot, q_lcl, q_icl, q_rai, q_sno, rho, N_rai, T
 *   out4 = HOST array of 4 device columns (NULL entries skipped): ∂ₜρn_rai, ∂ₜq_rai, ∂N_rai = ∂ₜρn_rai / N_rai (0 unless N_rai > eps),
 *          ∂q_rai = ∂ₜq_rai / q_rai (0 unless q_rai > eps) */
int cumicro_rain_evaporation_2m_f64(const cumicro_params_2m_warm_f64* p, int64_t n, const double* const* in8, double* const* out4, void* stream);
int cumicro_rain_evaporation_2m_f32(const cumicro_params_2m_warm_f32* p, int64_t n, const float* const* in8, float* const* out4, void* stream);

/* 2-moment terminal velocities (number- and mass-weighted).
 * CM2.rain_terminal_velocity(::SB2006, ::SB2006VelType, q_rai, rho, N_rai)  CM2:685-702
 * CM2.rain_terminal_velocity(::SB2006, ::Chen2022VelTypeRain, ...)          CM2:703-719
 * CM2.cloud_terminal_velocity(pdf_c, ::StokesRegimeVelType, q, rho, N)      CM2:647-664
 * N_* are number densities [1/m3] as in the reference signatures. */
int cumicro_termvel_2m_rain_sb_f64(const cumicro_sb_pdf_r_f64* pdf_r,
                                   const cumicro_vel_sb2006_f64* vel, int64_t n,
                                   const double* q_rai, const double* rho, const double* N_rai,
                                   double* vt0, double* vt1, void* stream);
int cumicro_termvel_2m_rain_chen_f64(const cumicro_sb_pdf_r_f64* pdf_r,
                                     const cumicro_vel_chen_rain_f64* vel, int64_t n,
                                     const double* q_rai, const double* rho, const double* N_rai,
                                     double* vt0, double* vt1, void* stream);
int cumicro_termvel_2m_cloud_f64(const cumicro_sb_pdf_c_f64* pdf_c,
                                 const cumicro_vel_stokes_f64* vel, int64_t n,
                                 const double* q_lcl, const double* rho, const double* N_lcl,
                                 double* vt0, double* vt1, void* stream);
int cumicro_termvel_2m_rain_sb_f32(const cumicro_sb_pdf_r_f32* pdf_r,
                                   const cumicro_vel_sb2006_f32* vel, int64_t n,
                                   const float* q_rai, const float* rho, const float* N_rai,
                                   float* vt0, float* vt1, void* stream);
int cumicro_termvel_2m_rain_chen_f32(const cumicro_sb_pdf_r_f32* pdf_r,
                                     const cumicro_vel_chen_rain_f32* vel, int64_t n,
                                     const float* q_rai, const float* rho, const float* N_rai,
                                     float* vt0, float* vt1, void* stream);
int cumicro_termvel_2m_cloud_f32(const cumicro_sb_pdf_c_f32* pdf_c,
                                 const cumicro_vel_stokes_f32* vel, int64_t n,
                                 const float* q_lcl, const float* rho, const float* N_lcl,
                                 float* vt0, float* vt1, void* stream);

/* ---------------------------------------------------------------------------
 * 1-moment scheme.  Input columns: rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno.
 * out4 = HOST array of 4 device column pointers: dq_lcl_dt, dq_icl_dt, dq_rai_dt, dq_sno_dt.
 *
 * cumicro_bmt1m_inst_*    replaces bulk_microphysics_tendencies(::Instantaneous,
 *                         ::Microphysics1Moment, mp, tps, ...)                    BMT:505-514
 * cumicro_bmt1m_verbose_* replaces the ::InstantaneousVerbose method              BMT:533-543
 *                         src18 = HOST array of 18 device column pointers (NULL entries skipped)
 *                         in the field order of _microphysics_source_terms       BMT:206-216:
 *                         S_phase_change_vap_lcl, S_phase_change_vap_icl, S_acnv_lcl_rai, S_acnv_icl_sno,
 *                         S_accr_lcl_rai, S_accr_lcl_sno_cold, S_accr_lcl_sno_warm, S_accr_melt_lcl_sno,
 *                         S_accr_icl_rai, S_accr_freeze_icl_rai, S_accr_icl_sno, S_accr_rai_sno_cold,
 *                         S_accr_rai_sno_warm, S_accr_melt_rai_sno, S_phase_change_vap_rai,
 *                         S_phase_change_vap_sno, S_melt_icl_lcl, S_melt_sno_rai
 * cumicro_bmt1m_linavg_*  replaces the ::LinearizedAverage method (dt, nsub)       BMT:572-632
 * ------------------------------------------------------------------------- */
#define CUMICRO_1M_NSRC 18
int cumicro_bmt1m_inst_f64(const cumicro_params_1m_f64* p, int64_t n, const double* rho, const double* T,
                           const double* q_tot, const double* q_lcl, const double* q_icl,
                           const double* q_rai, const double* q_sno, double* const* out4, void* stream);
int cumicro_bmt1m_inst_f32(const cumicro_params_1m_f32* p, int64_t n, const float* rho, const float* T,
                           const float* q_tot, const float* q_lcl, const float* q_icl,
                           const float* q_rai, const float* q_sno, float* const* out4, void* stream);
int cumicro_bmt1m_verbose_f64(const cumicro_params_1m_f64* p, int64_t n, const double* rho, const double* T,
                              const double* q_tot, const double* q_lcl, const double* q_icl,
                              const double* q_rai, const double* q_sno, double* const* out4,
                              double* const* src18, void* stream);
int cumicro_bmt1m_verbose_f32(const cumicro_params_1m_f32* p, int64_t n, const float* rho, const float* T,
                              const float* q_tot, const float* q_lcl, const float* q_icl,
                              const float* q_rai, const float* q_sno, float* const* out4,
                              float* const* src18, void* stream);
int cumicro_bmt1m_linavg_f64(const cumicro_params_1m_f64* p, int64_t n, const double* rho, const double* T,
                             const double* q_tot, const double* q_lcl, const double* q_icl,
                             const double* q_rai, const double* q_sno, double dt, int nsub,
                             double* const* out4, void* stream);
int cumicro_bmt1m_linavg_f32(const cumicro_params_1m_f32* p, int64_t n, const float* rho, const float* T,
                             const float* q_tot, const float* q_lcl, const float* q_icl,
                             const float* q_rai, const float* q_sno, float dt, int nsub,
                             float* const* out4, void* stream);

/* Terminal velocities of the 1-moment and non-equilibrium schemes, (rho, q) -> v [m/s].
 * kind 0: CM1.terminal_velocity(::Rain, ::Blk1MVelTypeRain, rho, q)              CM1:240-249
 *      1: CM1.terminal_velocity(::Snow, ::Blk1MVelTypeSnow, rho, q)
 *      2: CM1.terminal_velocity(::Rain, ::Chen2022VelTypeRain, rho, q)           CM1:251-270  (vel = cumicro_vel_chen_rain_*)
 *      3: CM1.terminal_velocity(::Snow, ::Chen2022VelTypeLargeIce, rho, q)       CM1:272-291  (vel = cumicro_vel_chen_large_ice_*)
 *      4: NEQ.terminal_velocity(::CloudLiquid, ::StokesRegimeVelType, rho, q)    NEQ:250-262  (vel = cumicro_vel_stokes_*)
 *      5: NEQ.terminal_velocity(::CloudIce, ::Chen2022VelTypeSmallIce, rho, q)   NEQ:264-281  (vel = cumicro_vel_chen_small_ice_*)
 * `vel` may be NULL for kinds 0 and 1 (the Blk1M parameters are part of `p`). */
int cumicro_termvel_1m_f64(const cumicro_params_1m_f64* p, const void* vel, int kind, int64_t n,
                           const double* rho, const double* q, double* out, void* stream);
int cumicro_termvel_1m_f32(const cumicro_params_1m_f32* p, const void* vel, int kind, int64_t n,
                           const float* rho, const float* q, float* out, void* stream);

/* ---------------------------------------------------------------------------
 * Ice nucleation, water activity (pointwise leaves): out[i] = fn(x[i] [, y[i]]).
 *  what 0: IN.deposition_J(dust, x = Δa_w)                       IN:92-102
 *       1: IN.ABIFM_J(dust, x = Δa_w)                            IN:124-134
 *       2: HomIceNucleation.homogeneous_J_cubic(koop, x = Δa_w)  IN:557-565
 *       3: HomIceNucleation.homogeneous_J_linear(koop, x = Δa_w) IN:581-584
 *       4: CO.a_w_ice(tps, x = T)                                CO:268-271
 *       5: CO.a_w_eT(tps, y = e, x = T)                          CO:256-258
 *       6: CO.a_w_xT(h2so4, tps, y = x_frac, x = T)              CO:241-245
 *       7: CO.H2SO4_soln_saturation_vapor_pressure(h2so4, y = x_frac, x = T)   CO:188-226
 *       8: IN.P3_deposition_N_i(mm2014, x = T)                   IN:162-166
 *       9: IN.INP_concentration_mean(frostenberg, x = T)         IN:250-253
 *      10: IN.dust_activated_number_fraction(dust, mohler, x = Si, y = T)      IN:44-52
 * `y` may be NULL for the one-argument functions.  Per-point domain violations (the reference
 * throws DomainError, IN:558-562, or fails an @assert, IN:47): the output is NaN and, if
 * `n_domain_errors` (a DEVICE counter the caller zeroes) is non-NULL, it is incremented.
 * ------------------------------------------------------------------------- */
int cumicro_icenuc_f64(const cumicro_params_icenuc_f64* p, int what, int64_t n, const double* x, const double* y,
                       double* out, unsigned long long* n_domain_errors, void* stream);
int cumicro_icenuc_f32(const cumicro_params_icenuc_f32* p, int what, int64_t n, const float* x, const float* y,
                       float* out, unsigned long long* n_domain_errors, void* stream);

/* Multi-argument nucleation rates: out[i] (, out2[i]) = fn(in5[0][i], ...); in5 = HOST array of device columns.
 *  what 0: IN.MohlerDepositionRate(dust, mohler, Si, T, dSi_dt, N_aer)                 IN:68-77   (4 columns)
 *          @assert Si < Sᵢ_max -> NaN + n_domain_errors
 *       1: IN.P3_het_N_i(mm2014, T, N_l, V_l, Δt)                                      IN:202-205 (4 columns)
 *       2: IN.INP_concentration_frequency(frostenberg, INPC, T)                        IN:219-224 (2 columns)
 *       3: P3.het_ice_nucleation(dust, tps, q_lcl, N_lcl, RH, T, ρₐ) -> out = dNdt, out2 = dLdt   P3_processes.jl:20-45 (5 columns)
 * out2 may be NULL. */
int cumicro_icenuc_rates_f64(const cumicro_params_icenuc_f64* p, int what, int64_t n, const double* const* in5, double* out,
                             double* out2, unsigned long long* n_domain_errors, void* stream);
int cumicro_icenuc_rates_f32(const cumicro_params_icenuc_f32* p, int what, int64_t n, const float* const* in5, float* out,
                             float* out2, unsigned long long* n_domain_errors, void* stream);

/* ARG2000 aerosol activation fused with the nucleation rates (BASELINE config 3).
 * Replaces AA.max_supersaturation (AA:138-214), AA.N_activated_per_mode (AA:235-273),
 * AA.M_activated_per_mode (AA:294-338) and IN.deposition_J / ABIFM_J / homogeneous_J_cubic
 * evaluated at Δa_w = CO.a_w_eT(tps, p_v, T) - CO.a_w_ice(tps, T), p_v the vapour pressure of
 * the same state — one read of the 8 state columns, all outputs written once.
 * N_act / M_act: HOST arrays of p->n_modes device column pointers (or NULL); any output
 * pointer may be NULL.  J_hom is NaN (+ counter) outside Koop's validity range. */
int cumicro_arg_icenuc_f64(const cumicro_params_icenuc_f64* p, int64_t n, const double* T, const double* p_air,
                           const double* w, const double* q_tot, const double* q_liq, const double* q_ice,
                           const double* N_liq, const double* N_ice, double* S_max, double* const* N_act,
                           double* const* M_act, double* J_dep, double* J_abifm, double* J_hom, double* da_w,
                           unsigned long long* n_domain_errors, void* stream);
int cumicro_arg_icenuc_f32(const cumicro_params_icenuc_f32* p, int64_t n, const float* T, const float* p_air,
                           const float* w, const float* q_tot, const float* q_liq, const float* q_ice,
                           const float* N_liq, const float* N_ice, float* S_max, float* const* N_act,
                           float* const* M_act, float* J_dep, float* J_abifm, float* J_hom, float* da_w,
                           unsigned long long* n_domain_errors, void* stream);

/* The trained-emulator methods AA.N_activated_per_mode(machine, ap, ad, aip, tps, T, p, w, qₜ, qₗ, qᵢ) and
 * AA.total_N_activated(machine, ...) of ext/EmulatorModelsExt.jl:32-103 for a multilayer-perceptron machine
 * (cumicro_params_emulator_*, cumicro_params.inc): N_act[i][k] = clamp(predict(row_i(k)), 0, 1) * mode_N[i].  The reference's
 * method ignores qₜ, qₗ, qᵢ as well.  `weights`: device buffer of cumicro_emulator_weight_count() values; N_act: HOST array of
 * n_modes device column pointers (any may be NULL); N_tot (may be NULL): their sum in mode order.  Dense layers run as 8x8x4
 * FP64 tensor-core matrix multiply-accumulates (32 rows per block); Float32 weights and columns are widened exactly and the
 * result is rounded once. */
int64_t cumicro_emulator_weight_count_f64(const cumicro_params_emulator_f64* p);
int64_t cumicro_emulator_weight_count_f32(const cumicro_params_emulator_f32* p);
int cumicro_aa_emulated_f64(const cumicro_params_emulator_f64* p, const double* weights, int64_t n, const double* T,
                            const double* p_air, const double* w, double* const* N_act, double* N_tot, void* stream);
int cumicro_aa_emulated_f32(const cumicro_params_emulator_f32* p, const float* weights, int64_t n, const float* T,
                            const float* p_air, const float* w, float* const* N_act, float* N_tot, void* stream);

/* ---------------------------------------------------------------------------
 * Fused 1-moment + 2-moment warm rain + ice nucleation (+ ARG2000 activation) with in-kernel
 * domain diagnostics (BASELINE config 5).  One read of the state, one write of every tendency.
 * in11  = HOST array of 11 device columns: rho, T, p, w, q_tot, q_lcl, q_icl, q_rai, q_sno, n_lcl, n_rai
 * out11 = HOST array of 11 device columns (NULL entries skipped):
 *         [0..3]  dq_lcl_dt, dq_icl_dt, dq_rai_dt, dq_sno_dt of the 1-moment scheme      BMT:505-514
 *         [4..7]  dq_lcl_dt, dn_lcl_dt, dq_rai_dt, dn_rai_dt of the 2-moment warm rain   BMT:820-854
 *                 (q_ice seen by its thermodynamics = q_icl + q_sno)
 *         [8..10] J_dep, J_ABIFM, J_hom at Δa_w = a_w_eT(p_v, T) - a_w_ice(T)              IN:92-134, 557-584
 * diag  = DEVICE array of CUMICRO_NDIAG doubles (or NULL), written (not accumulated):
 *         [0] Σ rho (dq_rai_dt + dq_sno_dt) 1-moment   [1] Σ rho dq_rai_dt 2-moment
 *         [2] Σ N_act (all modes of p3; 0 if p3->n_modes == 0)   [3] number of points
 *         Sums are Float64 and bit-reproducible (fixed reduction order).  Multi-GPU: each rank
 *         reduces its slab; the cross-rank sum is one all-reduce of CUMICRO_NDIAG doubles.
 * ------------------------------------------------------------------------- */
#define CUMICRO_NDIAG 4
int cumicro_fused_1m2m_icenuc_f64(const cumicro_params_1m_f64* p1, const cumicro_params_2m_warm_f64* p2,
                                  const cumicro_params_icenuc_f64* p3, int64_t n, const double* const* in11,
                                  double* const* out11, double* diag, void* stream);
int cumicro_fused_1m2m_icenuc_f32(const cumicro_params_1m_f32* p1, const cumicro_params_2m_warm_f32* p2,
                                  const cumicro_params_icenuc_f32* p3, int64_t n, const float* const* in11,
                                  float* const* out11, double* diag, void* stream);

/* ---------------------------------------------------------------------------
 * P3 ice scheme (src/P3_*.jl, src/Quadrature.jl).  Parameter block: cumicro_params_p3_* =
 * mp::Microphysics2MParams{WR, <:P3IceParams} + tps, flattened (quadrature nodes / weights of
 * mp.ice.quad included; the stand-alone reference functions take `quad` as a keyword).
 *
 * cumicro_p3_rates_*: the stand-alone P3 integrals of one state, BASELINE config 4.
 *   in12  = HOST array of 12 device columns: rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice,
 *           q_rim, b_rim, logλ  (specific quantities, as BMT takes them; volumetric L = q rho, N = n rho
 *           and state_from_prognostic are formed as at BMT:911-930; q_tot is not read)
 *   out12 = HOST array of 12 device columns (NULL entries skipped):
 *     [0] ice_terminal_velocity_number_weighted   [1] ..._mass_weighted      P3_terminal_velocity.jl:73-133
 *     [2] ice_melt dNdt  [3] ice_melt dLdt                                   P3_processes.jl:64-94
 *     [4] ice_self_collection dNdt                                           P3_processes.jl:676-712
 *     [5..11] bulk_liquid_ice_collision_sources: ∂ₜq_c, ∂ₜq_r, ∂ₜN_c, ∂ₜN_r, ∂ₜL_rim, ∂ₜL_ice, ∂ₜB_rim   :606-655
 *   Velocities are 0 where ρn_ice < eps or ρq_ice < eps (:79-81); [2..11] are 0 where the branch
 *   BMT:961 (q_ice > eps && n_ice > eps) is not taken.
 *
 * cumicro_bmt2m_p3_*: bulk_microphysics_tendencies(::Microphysics2Moment, mp{WR,<:P3IceParams}, tps, rho, T,
 *   q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, q_rim, b_rim, logλ, inpc_log_shift)       BMT:898-1083
 *   in12 as above; inpc_log_shift: device column or NULL (= 0);
 *   out9 = dq_lcl_dt, dn_lcl_dt, dq_rai_dt, dn_rai_dt, dq_ice_dt, dn_ice_dt, dq_rim_dt, db_rim_dt,
 *          dn_lcl_activation_dt (identically zero in the reference; NULL = not materialised).
 *
 * cumicro_termvel_p3_*: ice_terminal_velocity_{number,mass}_weighted_from_prognostic(vel, ρₐ, params, ρq_ice,
 *   ρn_ice, ρq_rim, ρb_rim, logλ; p = 1e-6, quad)                            P3_terminal_velocity.jl:135-173
 *
 * The logλ column of cumicro_p3_rates_* / cumicro_bmt2m_p3_* (in12[11]) and of cumicro_termvel_p3_* may be NULL: the kernel
 * then solves get_distribution_logλ_from_prognostic for every ice-bearing point itself (one point per thread, before the
 * integrals; the same code and bits as cumicro_p3_logl_* with brent_iters = 0) — SURVEY §8(f)-1.
 * cumicro_p3_logl_*: get_distribution_logλ_from_prognostic(params, ρq_ice, ρn_ice, ρq_rim, ρb_rim)
 *   P3_size_distribution.jl:284-334.  brent_iters <= 0 selects the reference's fixed 10 (Float64) / 8 (Float32)
 *   Brent iterations; -Inf for empty ice (:289).
 * ------------------------------------------------------------------------- */
int cumicro_p3_rates_f64(const cumicro_params_p3_f64* p, int64_t n, const double* const* in12, double* const* out12, void* stream);
int cumicro_p3_rates_f32(const cumicro_params_p3_f32* p, int64_t n, const float* const* in12, float* const* out12, void* stream);
int cumicro_bmt2m_p3_f64(const cumicro_params_p3_f64* p, int64_t n, const double* const* in12, const double* inpc_log_shift,
                         double* const* out9, void* stream);
int cumicro_bmt2m_p3_f32(const cumicro_params_p3_f32* p, int64_t n, const float* const* in12, const float* inpc_log_shift,
                         float* const* out9, void* stream);
int cumicro_termvel_p3_f64(const cumicro_params_p3_f64* p, int64_t n, const double* rho_a, const double* L_ice, const double* N_ice,
                           const double* L_rim, const double* B_rim, const double* logl, double* v_n, double* v_m, void* stream);
int cumicro_termvel_p3_f32(const cumicro_params_p3_f32* p, int64_t n, const float* rho_a, const float* L_ice, const float* N_ice,
                           const float* L_rim, const float* B_rim, const float* logl, float* v_n, float* v_m, void* stream);
int cumicro_p3_logl_f64(const cumicro_params_p3_f64* p, int64_t n, const double* L_ice, const double* N_ice, const double* L_rim,
                        const double* B_rim, int brent_iters, double* logl, void* stream);
int cumicro_p3_logl_f32(const cumicro_params_p3_f32* p, int64_t n, const float* L_ice, const float* N_ice, const float* L_rim,
                        const float* B_rim, int brent_iters, float* logl, void* stream);

/* ---------------------------------------------------------------------------
 * 0-moment scheme: bulk_microphysics_tendencies(::Microphysics0Moment, mp, tps, T, q_lcl, q_icl[, q_vap_sat])
 * BMT:658-680 -> CM0.remove_precipitation (src/Microphysics0M.jl:35-46):
 *   dq_tot_dt = -max(0, q_lcl + q_icl - threshold) / tau_precip, threshold = qc_0 (q_vap_sat == NULL) or S_0 q_vap_sat,
 * with the inputs clamped to >= 0 (BMT:662-663).  T does not enter the arithmetic and is not read.
 * ------------------------------------------------------------------------- */
int cumicro_bmt0m_f64(const cumicro_params_0m_f64* p, int64_t n, const double* q_lcl, const double* q_icl,
                      const double* q_vap_sat, double* dq_tot_dt, void* stream);
int cumicro_bmt0m_f32(const cumicro_params_0m_f32* p, int64_t n, const float* q_lcl, const float* q_icl,
                      const float* q_vap_sat, float* dq_tot_dt, void* stream);

/* The Frostenberg-2023 / Bigg nucleation rates of the 2-moment + P3 method on their own (BMT:998-1075):
 *   in9  = HOST array of 9 device columns: rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice (specific, clamped >= 0)
 *   inpc_log_shift: device column or NULL (= 0)
 *   out7 = HOST array of 7 device columns (NULL entries skipped):
 *     [0,1] IN.liquid_freezing_rate(rain_freezing, pdf_r, tps, q_rai, ρ, N_rai, T) -> ∂ₜn_frz, ∂ₜq_frz        IN:274-311
 *     [2,3] IN.liquid_freezing_rate(rain_freezing, pdf_c, tps, q_lcl, ρ, N_lcl, T) -> ∂ₜn_frz, ∂ₜq_frz        IN:356-388
 *     [4]   IN.immersion_limit_rate(frostenberg, T, ρ; τ = τ_act, inpc_log_shift, n_active = n_ice)         IN:420-430
 *     [5,6] IN.deposition_rate(frostenberg, tps, T, ρ, q_tot, q_lcl + q_rai, q_ice, n_ice; m_nuc, τ_act, inpc_log_shift)   IN:491-511 */
int cumicro_icenuc_f23_f64(const cumicro_params_p3_f64* p, int64_t n, const double* const* in9, const double* inpc_log_shift,
                           double* const* out7, void* stream);
int cumicro_icenuc_f23_f32(const cumicro_params_p3_f32* p, int64_t n, const float* const* in9, const float* inpc_log_shift,
                           float* const* out7, void* stream);

/* P3State thresholds and mass-weighted mean diameter over columns (volumetric inputs as state_from_prognostic takes them):
 *   out7 = F_rim, ρ_rim (regularised, clamped; P3_particle_properties.jl:101-106), ρ_g, D_th, D_gr, D_cr (:43-56; Inf when
 *   F_rim = 0), D_m(state, logλ) (P3_integral_properties.jl:56-61).  NULL entries are skipped. */
int cumicro_p3_state_f64(const cumicro_params_p3_f64* p, int64_t n, const double* L_ice, const double* N_ice, const double* L_rim,
                         const double* B_rim, const double* logl, double* const* out7, void* stream);
int cumicro_p3_state_f32(const cumicro_params_p3_f32* p, int64_t n, const float* L_ice, const float* N_ice, const float* L_rim,
                         const float* B_rim, const float* logl, float* const* out7, void* stream);
/* Shared numerics over columns (the reference tests them on the device, test/gpu_tests.jl:1305-1338): out = fn(x, y)
 *   what 0 / 1: UT.gamma_inc(a = x, x = y) -> P / Q  (UT:92-144; fixed 30 / 20 iterations)
 *        2: UT.gamma_inc_inv(a = x, p = y, q = 1 - y)  (UT:205-252)
 *        3: UT.rime_mass_fraction(q_rim = x, q_ice = y)   4: UT.rime_density(q_rim = x, b_rim = y)  (UT:445-509) */
int cumicro_p3_leaf_f64(int what, int64_t n, const double* x, const double* y, double* out, void* stream);
int cumicro_p3_leaf_f32(int what, int64_t n, const float* x, const float* y, float* out, void* stream);

/* ---------------------------------------------------------------------------
 * Alternative 2-moment closures (CM2:920-1002; goldens test/gpu_tests.jl:795-818): out = fn(columns)
 *   what 0: conv_q_lcl_to_q_rai(::KK2000, q_lcl, ρ, N_d)        1: (::B1994, ...; smooth_transition)
 *        2: (::TC1980, ...; smooth_transition)                 3: (::LD2004, ...; smooth_transition)
 *        4: accretion(::KK2000, q_lcl, q_rai, ρ)   5: accretion(::B1994, q_lcl, q_rai, ρ)   6: accretion(::TC1980, q_lcl, q_rai)
 *   Columns: q_lcl, q_rai (read by what >= 4, may be NULL otherwise), rho, N_d (read by what <= 3, may be NULL otherwise).
 * ------------------------------------------------------------------------- */
int cumicro_2m_alt_f64(const cumicro_params_2m_alt_f64* p, int what, int smooth_transition, int64_t n, const double* q_lcl,
                       const double* q_rai, const double* rho, const double* N_d, double* out, void* stream);
int cumicro_2m_alt_f32(const cumicro_params_2m_alt_f32* p, int what, int smooth_transition, int64_t n, const float* q_lcl,
                       const float* q_rai, const float* rho, const float* N_d, float* out, void* stream);

/* ---------------------------------------------------------------------------
 * Cloud diagnostics (src/CloudDiagnostics.jl:30-187; reference tests test/cloud_diagnostics.jl:30-125)
 *   cumicro_diag_2m_*:  radar_reflectivity_2M(sb, q_lcl, q_rai, N_lcl, N_rai, ρ_air) -> Z [dBZ] and
 *                       effective_radius_2M(...) -> r_eff [m] from one pass over the five columns (either output may be NULL)
 *   cumicro_diag_1m_*:  radar_reflectivity_1M(rain, q_rai, ρ_air) -> Z [dBZ]
 *   cumicro_diag_reff_lh97_*: effective_radius_Liu_Hallet_97((; ρw), ρ_air, q_lcl, N_lcl, q_rai, N_rai); N_lcl, q_rai, N_rai
 *                       NULL selects the three-argument method (N_lcl = 100, no rain).
 * ------------------------------------------------------------------------- */
int cumicro_diag_2m_f64(const cumicro_sb_pdf_c_f64* pdf_c, const cumicro_sb_pdf_r_f64* pdf_r, int64_t n, const double* q_lcl,
                        const double* q_rai, const double* N_lcl, const double* N_rai, const double* rho, double* Z, double* r_eff,
                        void* stream);
int cumicro_diag_2m_f32(const cumicro_sb_pdf_c_f32* pdf_c, const cumicro_sb_pdf_r_f32* pdf_r, int64_t n, const float* q_lcl,
                        const float* q_rai, const float* N_lcl, const float* N_rai, const float* rho, float* Z, float* r_eff,
                        void* stream);
int cumicro_diag_1m_f64(const cumicro_params_1m_f64* p, int64_t n, const double* q_rai, const double* rho, double* Z, void* stream);
int cumicro_diag_1m_f32(const cumicro_params_1m_f32* p, int64_t n, const float* q_rai, const float* rho, float* Z, void* stream);
int cumicro_diag_reff_lh97_f64(double rho_w, int64_t n, const double* rho, const double* q_lcl, const double* N_lcl, const double* q_rai,
                               const double* N_rai, double* r_eff, void* stream);
int cumicro_diag_reff_lh97_f32(float rho_w, int64_t n, const float* rho, const float* q_lcl, const float* N_lcl, const float* q_rai,
                               const float* N_rai, float* r_eff, void* stream);

/* ---------------------------------------------------------------------------
 * Multi-GPU: the grid is cut into independent column slabs, one per GPU (no halo, no collective on the tendency path).
 * The only exchange step is the optional domain diagnostics (BASELINE config 5):
 *   cumicro_reduce_diagnostics_*: one slab's sums out[k] = Σ_i weight[i] cols[k][i] (weight = rho, or NULL = 1) over `ncols`
 *     (<= 16) device columns (cols = HOST array of device pointers), Float64 accumulation in a fixed order (bit-reproducible).
 *     `out` = device array of ncols doubles; `scratch` = caller-owned device buffer of cumicro_reduce_diagnostics_scratch_bytes(ncols)
 *     bytes whose first 16 bytes are zero before the first use (the kernel leaves them zero); one scratch per stream in flight.
 *     (The fused config-5 kernel produces its CUMICRO_NDIAG sums in-kernel; this is the same reduction for the other entry points.)
 *   cumicro_nccl_allreduce_f64: in-place sum of `count` doubles over the ranks of `comm` (the caller's ncclComm_t), enqueued on
 *     `stream` (a side stream overlaps it with the next tendency kernel).  NCCL is resolved at run time from the process
 *     (dlsym) or libnccl.so.2; CUMICRO_E_NODEVICE if absent.
 *   cumicro_nccl_unique_id / comm_init_rank / comm_destroy: thin wrappers of ncclGetUniqueId / ncclCommInitRank /
 *     ncclCommDestroy (id128 = the 128-byte ncclUniqueId, exchanged by the host's own means) for hosts without an NCCL binding.
 * ------------------------------------------------------------------------- */
int64_t cumicro_reduce_diagnostics_scratch_bytes(int ncols);
int cumicro_reduce_diagnostics_f64(int64_t n, const double* weight, const double* const* cols, int ncols, double* out, void* scratch,
                                   int64_t scratch_bytes, void* stream);
int cumicro_reduce_diagnostics_f32(int64_t n, const float* weight, const float* const* cols, int ncols, double* out, void* scratch,
                                   int64_t scratch_bytes, void* stream);
int cumicro_nccl_allreduce_f64(void* comm, double* buf, int64_t count, void* stream);
int cumicro_nccl_unique_id(void* id128);
int cumicro_nccl_comm_init_rank(void** comm, int nranks, const void* id128, int rank);
int cumicro_nccl_comm_destroy(void* comm);

/* ---------------------------------------------------------------------------
 * The same exchange step without a library: peer-memory stores over NVLink / NVSwitch inside ONE kernel (csrc/cm_p2p.cuh).
 * Every rank (one process per GPU of one node) owns a window in its device memory; a rank writes its doubles into its slot of
 * every peer's window, waits for the peers' slots of its own window and adds them in rank order (bit-identical on all ranks).
 *   cumicro_p2p_window_create:  allocates and zeroes the local window on the current device (nranks <= 16); nranks = 1 needs no
 *                               connect step.
 *   cumicro_p2p_window_handle:  the CUMICRO_P2P_HANDLE_BYTES-byte inter-process handle of the local window (cudaIpcMemHandle_t);
 *                               the host ships it to the other ranks by its own means (MPI, torch.distributed, a file).
 *   cumicro_p2p_window_connect: `handles` = nranks x CUMICRO_P2P_HANDLE_BYTES bytes in rank order (the own entry is ignored);
 *                               maps the peers' windows.  The ranks must be separate processes with peer access between the GPUs.
 *   cumicro_p2p_allreduce_f64:  in-place sum of `count` (<= 16) doubles at device pointer `buf` over the ranks, one single-block
 *                               kernel on `stream`.  Every rank must make the same sequence of calls on its window.
 *   cumicro_fused_1m2m_icenuc_p2p_*: the config-5 entry point with the exchange inside its finish kernel: `diag` holds the
 *                               DOMAIN sums (all ranks) when the call's work completes — no second launch, no library call.
 *   cumicro_p2p_window_status:  calls made so far and the call number whose wait timed out (0 = none; a peer that never calls
 *                               makes the waiting ranks give up after the timeout, default 10 s, and return NaN sums).
 *   cumicro_p2p_window_destroy: after the last call's work has completed on EVERY rank (synchronise + barrier first).
 * ------------------------------------------------------------------------- */
#define CUMICRO_P2P_HANDLE_BYTES 64
int cumicro_p2p_window_create(int rank, int nranks, void** win);
int cumicro_p2p_window_handle(void* win, void* handle64);
int cumicro_p2p_window_connect(void* win, const void* handles);
int cumicro_p2p_window_set_timeout(void* win, double seconds);
int cumicro_p2p_window_status(void* win, int64_t* calls, int64_t* timed_out_call);
int cumicro_p2p_allreduce_f64(void* win, double* buf, int count, void* stream);
int cumicro_p2p_window_destroy(void* win);
int cumicro_fused_1m2m_icenuc_p2p_f64(const cumicro_params_1m_f64* p1, const cumicro_params_2m_warm_f64* p2,
                                      const cumicro_params_icenuc_f64* p3, int64_t n, const double* const* in11,
                                      double* const* out11, double* diag, void* win, void* stream);
int cumicro_fused_1m2m_icenuc_p2p_f32(const cumicro_params_1m_f32* p1, const cumicro_params_2m_warm_f32* p2,
                                      const cumicro_params_icenuc_f32* p3, int64_t n, const float* const* in11,
                                      float* const* out11, double* diag, void* win, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CUMICRO_H */
