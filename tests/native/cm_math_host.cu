// Host-side harness of cm_math.cuh for tests/test_cm_math.py: evaluates the HOST
// instantiation of the device math (same arithmetic, MUFU seeds emulated).
#include "../../cloudmicrophysics.jl_b200/csrc/cm_math.cuh"

extern "C" {
void cmt_exp(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = cm::exp_(x[i]); }
void cmt_exp_full(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = cm::exp_full_(x[i]); }
void cmt_log(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = cm::logp_(x[i]); }
void cmt_cbrt(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = cm::cbrtp_(x[i]); }
void cmt_rcp(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = cm::rcp_(x[i]); }
void cmt_sqrt(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = cm::sqrtp_(x[i]); }
// x / d through the shared, correctly rounded reciprocal (cm::divr_): y = quotient
void cmt_divr(const double* x, const double* d, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = cm::divr_(x[i], d[i], cm::rcp_cr_(d[i])); }
void cmt_log1p_pos(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = cm::log1p_pos_(x[i]); }
void cmt_rcbrt(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = cm::rcbrtp_(x[i]); }
void cmt_erf_fast(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = cm::erf_fast_(x[i]); }
void cmt_pow(const double* x, const double* p, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = cm::powp_(x[i], p[i]); }
}
