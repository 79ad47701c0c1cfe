// cm_math.cuh — math primitives of the cumicro kernels.
//
// The FP64 tendency kernels are bound by the FP64 pipe: B200
// issues 64 FP64 instructions/clk/SM, DADD/DMUL/DFMA alike, so the figure of merit
// of every function here is its FP64 *instruction count*.  The CUDA libm versions
// carry special-case handling and (for pow) double-double arithmetic that the call
// sites do not need; the versions below are written for the argument ranges the
// physics produces and stay far inside the 1e-12 relative parity budget:
//
//   exp_   : table-driven, 2^(j/256) from shared memory + degree-4 polynomial      10 FP64
//   logp_  : table-driven (128 x (1/c, -log 1/c)) + degree-7 log1p polynomial    ~13 FP64
//   powp_  : exp_(y * logp_(x)); relative error ~ |y ln x| * 3e-16               ~24 FP64
//   cbrtp_ : FP32 MUFU (lg2/ex2) seed, one cubic step on x^(-1/3), one correction ~12 FP64
//   rcp_   : MUFU.RCP64H seed + one cubically convergent step, <= 1 ulp              3 FP64
//   sqrtp_ : MUFU.RSQ64H seed + two coupled Newton steps, branch-free, < 1 ulp        7 FP64
// Suffix `p` = positive, normal, finite argument required (garbage, not NaN, outside;
// call sites select the result away in those cases, exactly where the reference's
// ifelse does).  + - * / sqrt are IEEE operations, identical to the CPU reference; the
// code is compiled with -fmad=false, so a fused multiply-add happens exactly where
// fma() is written and results do not depend on how the compiler schedules the
// surrounding code (scalar-tail and vector kernels give identical bits).
//
// Every function is __host__ __device__ (MUFU seeds are emulated on the host) so
// tests/test_cm_math.py measures the accuracy against mpmath without a GPU.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cmath>
#include <cstdint>
#include <cstring>

#include "cm_math_tables.inc"

namespace cm {

#define CM_DEV __device__ __forceinline__
#define CM_HD __host__ __device__ __forceinline__

// ---- bit access -------------------------------------------------------------------------
CM_HD int hi32(double x) {
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    int64_t u; memcpy(&u, &x, 8); return (int)(u >> 32);
#endif
}
CM_HD int lo32(double x) {
#ifdef __CUDA_ARCH__
    return __double2loint(x);
#else
    int64_t u; memcpy(&u, &x, 8); return (int)(u & 0xffffffff);
#endif
}
CM_HD double mk64(int hi, int lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &u, 8); return d;
#endif
}
CM_HD double bits2d(unsigned long long u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}

template <class FT> struct num;
template <> struct num<double> {
    static CM_HD double eps() { return 2.220446049250313e-16; }
    // UT.ϵ_numerics(Float64) = cbrt(floatmin(Float64))            UT:318
    static CM_HD double eps_numerics() { return 2.8126442852362996e-103; }
    static CM_HD double inf() { return bits2d(0x7ff0000000000000ULL); }
    static CM_HD double pi() { return 3.141592653589793; }
};
template <> struct num<float> {
    static CM_HD float eps() { return 1.1920929e-07f; }
    // cbrt(floatmin(Float32))
    static CM_HD float eps_numerics() { return 2.2737368e-13f; }
    static CM_HD float inf() { return (float)bits2d(0x7ff0000000000000ULL); }
    static CM_HD float pi() { return 3.1415927f; }
};

// ---- lookup tables ------------------------------------------------------------------------
// Global copies (host + device) and the per-block shared-memory copy the device
// functions read.  A kernel that uses exp_/logp_/powp_ must call math_tables_init()
// (all threads of the block) before its first use; cm_launch.cuh does.
static const unsigned long long cm_exp_tab_host[256] = CM_EXP_TABLE_INIT;
static const unsigned long long cm_log_tab_host[256] = CM_LOG_TABLE_INIT;
static __device__ const unsigned long long cm_exp_tab_dev[256] = CM_EXP_TABLE_INIT;
static __device__ const unsigned long long cm_log_tab_dev[256] = CM_LOG_TABLE_INIT;
static const unsigned long long cm_log2_tab_host[512] = CM_LOG2_TABLE_INIT;
static __device__ const unsigned long long cm_log2_tab_dev[512] = CM_LOG2_TABLE_INIT;
#ifdef __CUDACC__
static __shared__ unsigned long long cm_sh_exp[256];
static __shared__ ulonglong2 cm_sh_log[128];
static __shared__ ulonglong2 cm_sh_log2[256];   // log_abs_ (allocated only in kernels that call math_tables_init_log2)
// the few coefficients that need all 53 bits (everything else in exp_/logp_/cbrtp_ is an
// immediate operand: constants whose low 32 bits are zero cost no instruction on sm_100)
static __constant__ double cm_kc[8] = {3.3333333333333331e-01, 0.2, CM_LOG_LN2_LO, -(CM_EXP_L2F), CM_EXP_INV_L, CM_EXP_C4, CM_LN2, 0.0};
#endif
#define CM_LOG_C3_HOST 3.3333333333333331e-01

// WITH_LOG = false: a kernel that never calls logp_ (the 2M tile kernel uses log_abs_) leaves the 2 KB logp_ table out
template <int BLOCK, bool WITH_LOG = true> CM_DEV void math_tables_init() {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int i = threadIdx.x; i < 256; i += BLOCK) cm_sh_exp[i] = cm_exp_tab_dev[i];
    if (WITH_LOG) {
#pragma unroll
        for (int i = threadIdx.x; i < 128; i += BLOCK)
            cm_sh_log[i] = make_ulonglong2(cm_log_tab_dev[2 * i], cm_log_tab_dev[2 * i + 1]);
    }
    __syncthreads();
#endif
}
template <int BLOCK> CM_DEV void math_tables_init_log2() {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int i = threadIdx.x; i < 256; i += BLOCK)
        cm_sh_log2[i] = make_ulonglong2(cm_log2_tab_dev[2 * i], cm_log2_tab_dev[2 * i + 1]);
    __syncthreads();
#endif
}
CM_HD void log2_tab(int i, double& invc, double& logc) {
#ifdef __CUDA_ARCH__
    const ulonglong2 t = cm_sh_log2[i];
    invc = bits2d(t.x);
    logc = bits2d(t.y);
#else
    invc = bits2d(cm_log2_tab_host[2 * i]);
    logc = bits2d(cm_log2_tab_host[2 * i + 1]);
#endif
}
CM_HD double exp_tab(int j) {
#ifdef __CUDA_ARCH__
    return bits2d(cm_sh_exp[j]);
#else
    return bits2d(cm_exp_tab_host[j]);
#endif
}
CM_HD void log_tab(int i, double& invc, double& logc) {
#ifdef __CUDA_ARCH__
    const ulonglong2 t = cm_sh_log[i];
    invc = bits2d(t.x);
    logc = bits2d(t.y);
#else
    invc = bits2d(cm_log_tab_host[2 * i]);
    logc = bits2d(cm_log_tab_host[2 * i + 1]);
#endif
}

// ---- min / max / clamp with the reference's (Julia Base) selection semantics --------------
// Base.max / Base.min as one compare + select ((a < b) ? b : a, the oracle's jmax): IEEE fmax()/fmin() cost
// ~7 instructions per call on sm_100 (NaN quieting) and do not have Julia's semantics either.
CM_HD double fmax_(double a, double b) { return (a < b) ? b : a; }
CM_HD float fmax_(float a, float b) { return (a < b) ? b : a; }
CM_HD double fmin_(double a, double b) { return (b < a) ? b : a; }
CM_HD float fmin_(float a, float b) { return (b < a) ? b : a; }
// max(0, x) / min(0, x): the compiler recognises (0 < x) ? x : 0 as an IEEE maxnum and expands it into DSETP.MAX + selects + a
// NaN-quieting LOP3 (6-7 instructions, two of them on the FP64 pipe); the sign-bit form is three integer instructions.
// -0.0 and negative NaNs clamp to +0.0; a positive NaN propagates (as in Julia's max).
CM_HD double clamp0_(double x) {
#ifdef __CUDA_ARCH__
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, lo & keep);
#else
    return std::signbit(x) ? 0.0 : x;
#endif
}
CM_HD double cap0_(double x) {
#ifdef __CUDA_ARCH__
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int keep = hi >> 31;
    return __hiloint2double(hi & keep, lo & keep);
#else
    return std::signbit(x) ? x : 0.0;
#endif
}
CM_HD float clamp0_(float x) { return (0.0f < x) ? x : 0.0f; }
CM_HD float cap0_(float x) { return (x < 0.0f) ? x : 0.0f; }
// Base.clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x))
template <class FT> CM_HD FT clamp_(FT x, FT lo, FT hi) { return (x > hi) ? hi : ((x < lo) ? lo : x); }
CM_HD double fma_(double a, double b, double c) { return fma(a, b, c); }
CM_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }

// ---- IEEE operations (bit-identical to the CPU reference) ------------------------------------
CM_HD double sqrt_(double x) { return sqrt(x); }
CM_HD float sqrt_(float x) { return sqrtf(x); }
CM_HD double div_(double a, double b) { return a / b; }
CM_HD float div_(float a, float b) { return a / b; }

// ---- reciprocal: <= 1 ulp, normal finite x ---------------------------------------------------
CM_HD double rcp_(double x) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));  // MUFU.RCP64H: ~20 good bits
#else
    double r = mk64(hi32(1.0 / x), 0);  // host emulation of the 20-bit seed
#endif
    // one cubically convergent step: e = 1 - x r (2^-20), r (1 + e + e^2) is good to e^3 = 2^-60
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}
// ---- IEEE-quality division through a shared reciprocal ----------------------------------------
// CUDA's a / b leaves its ~13-instruction fast path for a ~100-instruction subroutine whenever the NUMERATOR is zero (or below
// 2^-969) — and gated-off source terms are exact zeros.  Where several quotients share a divisor (the four max(q_min, q) of
// BMT._linearize, the two determinants of the 2x2 solves) one reciprocal r = RN(1/d) serves all of them:
//   q = RN(x r), rem = x - q d (exact, FMA), q' = RN(q + rem r) is the correctly rounded x / d (Markstein 1990, Thm 3.1;
//   r itself is correctly rounded after one FMA correction of the <= 1 ulp rcp_, except for divisors with an all-ones
//   significand, where the quotient can be 1 ulp off).  x = 0 gives 0 (the sign of a zero quotient is not preserved), no branch.
// d must be positive, normal and finite.
CM_HD double rcp_cr_(double d) {
    const double r = rcp_(d);
    return fma(r, fma(-d, r, 1.0), r);
}
CM_HD double divr_(double x, double d, double r) {
    const double q = x * r;
    return fma(fma(-q, d, x), r, q);
}
CM_HD float rcp_cr_(float d) { return 1.0f / d; }
CM_HD float divr_(float x, float d, float) { return x / d; }

// ---- square root: positive, normal, finite x; faithfully rounded (< 1 ulp), no special cases -------
CM_HD double sqrtp_(double x) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));  // MUFU.RSQ64H: ~20 good bits
#else
    double r = mk64(hi32(1.0 / std::sqrt(x)), 0);
#endif
    double g = x * r;          // ~ sqrt(x)
    double h = 0.5 * r;        // ~ 1 / (2 sqrt(x))
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);          // 40 bits
    e = fma(-h, g, 0.5);
    return fma(g, e, g);       // 80 bits -> rounding only
}
CM_HD float sqrtp_(float x) { return sqrtf(x); }
// guarded form: the faithful 7-instruction sqrtp_ for positive normal arguments, IEEE sqrt (0, subnormal, Inf, NaN, negative) otherwise
CM_HD double sqrtg_(double x) { return (x > 2.3e-308 && x < 1.7e308) ? sqrtp_(x) : sqrt(x); }
CM_HD float sqrtg_(float x) { return sqrtf(x); }
// 1 / sqrt(x): the same coupled iteration read out on its other variable (9 FP64 instructions, < 1.5 ulp) for positive normal
// arguments, IEEE 1 / sqrt(x) otherwise
CM_HD double rsqrtg_(double x) {
    if (!(x > 2.3e-308 && x < 1.7e308)) return 1.0 / sqrt(x);
#ifdef __CUDA_ARCH__
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#else
    double r = mk64(hi32(1.0 / std::sqrt(x)), 0);
#endif
    double g = x * r;
    double h = 0.5 * r;
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-h, g, 0.5);
    h = fma(h, e, h);
    return h + h;
}
// x^(1/3) for x >= 0: the 12-instruction cbrtp_ for positive normal arguments, the CUDA libm's otherwise (0, subnormal, Inf, NaN)
CM_HD double cbrt_pair_(double x, double& rc);
CM_HD double cbrtg_(double x) {
    if (!(x > 2.3e-308 && x < 1.7e308)) return cbrt(x);
    double rc;
    return cbrt_pair_(x, rc);
}
CM_HD float rcp_(float x) { return 1.0f / x; }

// ---- exp: |x| <= 708 (callers' arguments are bounded; see exp_full_ otherwise) ---------------
// x = (256 e + j) ln2/256 + r, |r| <= ln2/512;  exp(x) = 2^e * T[j] * (1 + expm1(r)).
// 9 FP64 instructions, 1 LDS, ~5 integer: ln2/256 = L1 (21 bits, immediate, kf L1 exact) + L2F (full double, constant bank).
#ifndef CM_EXP_ROT
#define CM_EXP_ROT 1   /* 2M headline 0.3915 -> 0.3905 ms, P3 48.35 -> 48.18 ms per 2^20 points; same bits */
#endif
#ifndef CM_EXP_LEA
#define CM_EXP_LEA 1   /* 2M headline 0.3994 -> 0.3920 ms, config 3 1.454 -> 1.442 ms, config 5 1.845 -> 1.839 ms; same bits */
#endif
CM_HD double exp_(double x) {
    const double magic = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer
    // an FP64 instruction takes ONE non-register operand: the second constant of a two-constant fma comes from the
    // constant bank (one uniform load) instead of two 32-bit moves
#ifdef __CUDA_ARCH__
    const double inv_l = cm_kc[4], c4 = cm_kc[5], nl2f = cm_kc[3];
#else
    const double inv_l = CM_EXP_INV_L, c4 = CM_EXP_C4, nl2f = -(CM_EXP_L2F);
#endif
    const double t = fma(x, inv_l, magic);
    const int ki = lo32(t);
    const double kf = t - magic;
    double r = fma(kf, -CM_EXP_L1, x);  // exact
    r = fma(kf, nl2f, r);
    double p = fma(r, c4, CM_EXP_C3);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = p * r;  // expm1(r)
#if defined(__CUDA_ARCH__) && CM_EXP_ROT
    unsigned off;   // byte offset of T[ki & 255]: a funnel shift (rotate by 3) + mask stay on the ALU pipe; `<< 3` becomes IMAD.SHL
    asm("{.reg .b32 t; shf.l.wrap.b32 t, %1, %1, 3; and.b32 %0, t, 0x7f8;}" : "=r"(off) : "r"(ki));
    const double T = bits2d(*reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(cm_sh_exp) + off));
#else
    const double T = exp_tab(ki & 255);
#endif
    const double y = fma(T, p, T);
#if defined(__CUDA_ARCH__) && CM_EXP_LEA
    // 2^(ki >> 8) onto the high word as an arithmetic shift, a shift and an add written in PTX: ptxas keeps them on the ALU pipe
    // (SHF), while it turns the C expression below into IMAD.SHL + LOP3 + IMAD.IADD — and IMAD shares its issue path with the
    // FP64 instructions this body is made of (the same integer result; measured, see CM_EXP_LEA)
    int hi2;
    asm("{.reg .s32 t; shr.s32 t, %1, 8; shl.b32 t, t, 20; add.s32 %0, t, %2;}" : "=r"(hi2) : "r"(ki), "r"(hi32(y)));
    return mk64(hi2, lo32(y));
#else
    return mk64(hi32(y) + (ki & ~255) * 4096, lo32(y));   // 2^(ki >> 8): one mask + one multiply-add on the high word
#endif
}
// exp with the IEEE limits: gradual underflow into the subnormals, 0 below them, +Inf
// above the range, NaN propagated.
CM_HD double exp_full_(double x) {
    // In range (all but pathological points, and warp-uniformly so) this is one compare, exp_ itself and one skipped branch;
    // the branch-free form (six selects, IEEE fmin/fmax, three compares) cost more than the exponential — ncu source view of
    // the ARG2000 kernel: 20 % of its instructions.  ONE inlined copy of exp_ serves both paths.
    // Out of range: one exp_ evaluation on a shifted argument: e^x = e^(x + 64 ln2) 2^-64 below the normal range (the scaling
    // multiply rounds once into the subnormals), e^x = e^(x - 1) e just below overflow.
    const bool special = !(fabs(x) <= 708.0);   // also NaN
    double xs = x, sc = 1.0;
    if (special) {
        const bool lo = x < -708.0, hi = x > 709.0;
        xs = lo ? x + 44.361419555836500 : (hi ? x - 1.0 : x);
        sc = lo ? 5.421010862427522170e-20 : (hi ? 2.718281828459045 : 1.0);
        xs = fmin(fmax(xs, -708.0), 709.0);
    }
    double y = exp_(xs);
    if (special) {
        y *= sc;
        y = (x < -746.0) ? 0.0 : y;
        y = (x > 709.782712893384) ? num<double>::inf() : y;
        y = (x != x) ? x : y;
    }
    return y;
}
CM_HD float exp_(float x) { return expf(x); }
CM_HD float exp_full_(float x) { return expf(x); }

// ---- log: positive, normal, finite x ------------------------------------------------------
// x = 2^k z, z in [0.6875, 1.375); z = c (1 + r) with 1/c, -log(1/c) tabulated (128 cells).
CM_HD double logp_(double x) {
#ifdef __CUDA_ARCH__
    const double c3 = cm_kc[0], c5 = cm_kc[1], ln2_lo = cm_kc[2];
#else
    const double c3 = CM_LOG_C3_HOST, c5 = 0.2, ln2_lo = CM_LOG_LN2_LO;
#endif
    const int hx = hi32(x);
    const int tmp = hx - 0x3fe60000;
    const int i = (tmp >> 13) & 127;
    const int k = tmp >> 20;  // arithmetic shift: floor
    const double z = mk64(hx - (k << 20), lo32(x));
    double invc, logc;
    log_tab(i, invc, logc);
    const double r = fma(z, invc, -1.0);
    const double kd = (double)k;
    const double w = fma(kd, CM_LOG_LN2_HI, logc);  // k * LN2_HI is exact (21-bit constant)
    double q = fma(r, CM_LOG_C7, CM_LOG_C6);
    q = fma(q, r, c5);
    q = fma(q, r, -0.25);
    q = fma(q, r, c3);
    q = fma(q, r, -0.5);
    const double r2 = r * r;
    const double lo = fma(r2, q, kd * ln2_lo);
    return (w + r) + lo;
}
CM_HD float logp_(float x) { return logf(x); }
// log with ABSOLUTE accuracy ~1.5e-16 max(1, |log x|) (not relatively accurate near x = 1): positive, normal, finite x.
// x = 2^k z, z in [1, 2); z = c (1 + r) with 1/c, -log(1/c) tabulated (256 cells, |r| <= 2^-9); degree-5 log1p: 8 FP64.
// For arguments whose logarithm is an additive term of an exponent or of a log-space recurrence (log L, log N, log tau).
CM_HD double log_abs_(double x) {
#ifdef __CUDA_ARCH__
    const double c3 = cm_kc[0], ln2 = cm_kc[6];
#else
    const double c3 = CM_LOG_C3_HOST, ln2 = CM_LN2;
#endif
    const int hx = hi32(x);
    const int i = (hx >> 12) & 255;
    const int k = (hx >> 20) - 1023;
    const double z = mk64((hx & 0x000fffff) | 0x3ff00000, lo32(x));
    double invc, logc;
    log2_tab(i, invc, logc);
    const double r = fma(z, invc, -1.0);
    const double w = fma((double)k, ln2, logc);
    double q = fma(r, CM_LOGA_C5, -0.25);
    q = fma(q, r, c3);
    q = fma(q, r, -0.5);
    return w + fma(r * r, q, r);
}
CM_HD float log_abs_(float x) { return logf(x); }
CM_HD double log_full_(double x) { return log(x); }
CM_HD float log_full_(float x) { return logf(x); }

// ---- log1p for a positive finite argument (the e^x of log1pexp): log(u) y / (u - 1) with u = 1 + y (Kahan) --------------
// u - 1 is exact (Sterbenz / exponent alignment), so the quotient y/(u-1) carries the rounding of 1 + y back out: ~4 ulp
// for every y > 0 with the table-driven logarithm, against ~50 FP64 instructions of the libm log1p.
CM_HD double log1p_pos_(double y) {
    const double u = 1.0 + y;
    const double d = u - 1.0;
    return (d == 0.0) ? y : logp_(u) * (y * rcp_(d));
}
CM_HD float log1p_pos_(float y) { return log1pf(y); }

// ---- pow: x positive normal, |y ln x| <= 708 ---------------------------------------------
CM_HD double powp_(double x, double y) { return exp_(y * logp_(x)); }
CM_HD float powp_(float x, float y) { return powf(x, y); }
CM_HD double pow_full_(double x, double y) { return pow(x, y); }
CM_HD float pow_full_(float x, float y) { return powf(x, y); }

// ---- cbrt: positive, normal, finite x -------------------------------------------------------
// cbrt_pair_: y = x^(1/3) (< 1 ulp) and rc = x^(-1/3) (< 1.5 ulp) for two more FMAs: the iterate the cube root is built from, refined
CM_HD double cbrt_pair_(double x, double& rc) {
    const int hx = hi32(x);
    const int e = (hx >> 20) - 1023;               // x = m 2^e, m in [1, 2)
    const int q = ((e + 3072) * 43691 >> 17) - 1024;  // floor(e / 3) for |e| <= 1100
    const double a = mk64(hx - ((3 * q) << 20), lo32(x));  // a = x 2^(-3q) in [1, 8)
    // r0 ~ a^(-1/3) from the FP32 special-function unit (relative error ~ 2^-21)
#ifdef __CUDA_ARCH__
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"((float)a));
    float rf;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(l * -0.33333334f));
    double r = (double)rf;
#else
    double r = (double)(float)(1.0 / std::cbrt(a)) * (1.0 + 3e-7);
#endif
    // one cubically convergent step on r -> a^(-1/3): e = 1 - a r^3, r *= 1 + e/3 + 2 e^2/9
    double r2 = r * r;
    const double err = fma(-(a * r), r2, 1.0);
    const double pe = fma(err, CM_CBRT_C2, CM_CBRT_C1) * err;
    r = fma(r, pe, r);
    r2 = r * r;
    double y = a * r2;  // a^(1/3), ~2-3 ulp
    // one Newton correction with the residual computed by fma: y -= (y^3 - a) / (3 y^2)
    const double d = fma(-(y * y), y, a);
    y = fma(d, r2 * CM_CBRT_C1, y);
    // r is a^(-1/3) to ~1e-13 only (the 21-bit immediate 1/3 above): one Newton step against the finished cube root
    const double rn = fma(r, fma(-y, r, 1.0), r);
    rc = mk64(hi32(rn) - (q << 20), lo32(rn));
    return mk64(hi32(y) + (q << 20), lo32(y));
}
CM_HD double cbrtp_(double x) {
    double rc;
    return cbrt_pair_(x, rc);
}
CM_HD double rcbrtp_(double x) {   // x^(-1/3), ~2 ulp
    double rc;
    cbrt_pair_(x, rc);
    return rc;
}
CM_HD float cbrt_pair_(float x, float& rc) { const float y = cbrtf(x); rc = 1.0f / y; return y; }
CM_HD float rcbrtp_(float x) { return 1.0f / cbrtf(x); }
CM_HD float cbrtp_(float x) { return cbrtf(x); }
CM_HD double cbrt_full_(double x) { return cbrt(x); }
CM_HD float cbrt_full_(float x) { return cbrtf(x); }

// ---- rarely used special functions: CUDA libm ------------------------------------------------
CM_HD double expm1_(double x) { return expm1(x); }
CM_HD float expm1_(float x) { return expm1f(x); }
CM_HD double log1p_(double x) { return log1p(x); }
CM_HD float log1p_(float x) { return log1pf(x); }
CM_HD double tgamma_(double x) { return tgamma(x); }
CM_HD float tgamma_(float x) { return tgammaf(x); }
CM_HD double lgamma_(double x) { return lgamma(x); }
CM_HD float lgamma_(float x) { return lgammaf(x); }
// erf_ is the CUDA libm's — on sm_100a itself ONE branch-free piece (41 FP64 + a MUFU.EX2 + 68 UMOVs that ptxas hoists out of a
// grid-stride loop).  Replacements were measured three times: round 1, a table-driven one with its polynomials in shared memory
// (config 3 2.01 -> 2.08 ms: lanes of a warp sit in different pieces, 15 dependent LDS + DFMA pairs): dropped; round 2, erf_fast_
// below in the grid-stride ARG2000 kernel (2.01 -> 1.99 ms) and in the fused kernel (2.36 -> 2.42 ms): no gain, because the
// libm's constant moves sat outside those loops; then in the TILE-shaped ARG2000 kernel, whose loop does not hoist them:
// 1.68 -> 1.57 ms — kept there (arg2000<WANT_M, FAST_ERF = true>), while the fused kernel stays on the libm's (2.22 vs 2.31 ms).
// erf_fast_: ONE branch-free piece, erf(|x|) = 1 - exp_(-|x| Q(t)), t = |x|/4 - 3/4, |x| clamped to 6 (erfc(6) = 2e-17: the result
// is 1.0), Q = -log(erfc(x)) / x as a degree-20 polynomial whose coefficients come from the constant bank two per uniform load
// (tools/gen_math_tables.py: weighted fit of the ABSOLUTE error of erf).  |error| <= 2.5 units of 2^-53 ABSOLUTE — not relative
// for |x| -> 0, which its use, N (1 - erf(u)) / 2 (AA:256), does not need.  35 FP64 + ~35 other instructions, all inside the
// loop; the libm's 68 UMOVs are only free where ptxas can hoist them (a grid-stride loop), not in the tile loop: the ARG2000
// kernel (tile shape) runs this one (config 3: 1.68 -> 1.57 ms), the fused kernel keeps the libm's (2.22 vs 2.31 ms with this).
#ifdef __CUDACC__
static __constant__ __align__(16) double cm_erf_q[CM_ERF_DEG + 1 + ((CM_ERF_DEG + 1) & 1)] = CM_ERF_Q_INIT;
#endif
static const double cm_erf_q_host[CM_ERF_DEG + 1] = CM_ERF_Q_INIT;
CM_HD double exp_(double x);
CM_HD double erf_fast_(double x) {
#ifdef __CUDA_ARCH__
    const double* c = cm_erf_q;
#else
    const double* c = cm_erf_q_host;
#endif
    const double a0 = fabs(x);
    const double a = (a0 > 6.0) ? 6.0 : a0;   // NaN stays
    const double t = fma(a, 0.25, -0.75);
    double q = c[CM_ERF_DEG];
#pragma unroll
    for (int i = CM_ERF_DEG - 1; i >= 0; --i) q = fma(q, t, c[i]);
    const double r = 1.0 - exp_(-(a * q));     // a q in [0, 38.4]
    const double y = copysign(r, x);
    return (x != x) ? x : y;
}
CM_HD double erf_(double x) { return erf(x); }
CM_HD float erf_(float x) { return erff(x); }
CM_HD double erfc_(double x) { return erfc(x); }
CM_HD float erfc_(float x) { return erfcf(x); }
CM_HD double tanh_(double x) { return tanh(x); }
CM_HD float tanh_(float x) { return tanhf(x); }

// x^y for a parameter exponent that is very often a small integer (SB2006's
// b = 3, c = 4, d = -5): the integer cases are exact products (what Julia's ^ gives for
// these literal-valued parameters to within 1 ulp), everything else goes through powp_.
// `y` is uniform across the grid, so the branch is free.
template <class FT> CM_HD FT pow_param(FT x, FT y) {
    if (y == FT(2)) return x * x;
    if (y == FT(3)) return x * x * x;
    if (y == FT(4)) { FT x2 = x * x; return x2 * x2; }
    if (y == FT(-5)) { FT x2 = x * x; return FT(1) / (x2 * x2 * x); }
    if (y == FT(1)) return x;
    return powp_(x, y);
}
// Compile-time integer power: exact products; N < 0 through the <= 1 ulp reciprocal (x positive, normal).
template <int N, class FT> CM_HD FT pow_int_(FT x) {
    if constexpr (N < 0) return rcp_(pow_int_<-N>(x));
    else if constexpr (N == 0) return FT(1);
    else if constexpr (N == 1) return x;
    else if constexpr (N % 2 == 0) { const FT h = pow_int_<N / 2>(x); return h * h; }
    else return pow_int_<N - 1>(x) * x;
}
// The same with the case decided once on the host (an integer code in the launch constants) instead of up to
// five FP64 compares per point.
template <class FT> inline int pow_param_code(FT y) {
    return (y == FT(1)) ? 1 : (y == FT(2)) ? 2 : (y == FT(3)) ? 3 : (y == FT(4)) ? 4 : (y == FT(-5)) ? 5 : 0;
}
template <class FT> CM_HD FT pow_param(FT x, FT y, int code) {
    switch (code) {
        case 1: return x;
        case 2: return x * x;
        case 3: return x * x * x;
        case 4: { FT x2 = x * x; return x2 * x2; }
        case 5: { FT x2 = x * x; return rcp_(x2 * x2 * x); }   // x positive normal at every call site; same bits as pow_int_<-5>
        default: return powp_(x, y);
    }
}

// --- 128-bit vector access ------------------------------------------------------------
template <class FT> struct vec;
template <> struct vec<double> {
    using type = double2;
    static constexpr int N = 2;
};
template <> struct vec<float> {
    using type = float4;
    static constexpr int N = 4;
};

template <class FT, int N> struct pack { FT v[N]; };

#ifdef __CUDACC__
CM_DEV pack<double, 2> ldg_vec(const double* p) {
    double2 t = __ldg(reinterpret_cast<const double2*>(p));
    pack<double, 2> r; r.v[0] = t.x; r.v[1] = t.y; return r;
}
CM_DEV pack<float, 4> ldg_vec(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    pack<float, 4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
CM_DEV void st_vec(double* p, const pack<double, 2>& r) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(r.v[0], r.v[1]));
}
CM_DEV void st_vec(float* p, const pack<float, 4>& r) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
}
#endif

}  // namespace cm
