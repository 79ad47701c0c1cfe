// oracle_api.cpp — extern "C" array drivers over the scalar CPU restatement.
// TEST INFRASTRUCTURE ONLY (see oracle_base.hpp).  OpenMP over points: this is
// also the "port" CPU baseline timed by bench.py (the reference has no
// multithreaded driver of its own; hosts broadcast the scalar methods).
#include <omp.h>

#include "oracle_2m.hpp"

using namespace orc;

extern "C" {

int oracle_num_threads(void) { return omp_get_max_threads(); }
void oracle_set_num_threads(int n) { omp_set_num_threads(n); }

#define DEF_BMT2M_WARM(SUF, FT)                                                                      \
    int oracle_bmt2m_warm_##SUF(const cumicro_params_2m_warm_##SUF* p, int64_t n, const FT* rho,     \
                                const FT* T, const FT* q_tot, const FT* q_lcl, const FT* n_lcl,      \
                                const FT* q_rai, const FT* n_rai, FT* dq_lcl_dt, FT* dn_lcl_dt,      \
                                FT* dq_rai_dt, FT* dn_rai_dt, FT* const* leaves) {                   \
        _Pragma("omp parallel for schedule(static)") for (int64_t i = 0; i < n; ++i) {               \
            Warm2MOut<FT> o = bmt2m_warm<FT>(*p, rho[i], T[i], q_tot[i], q_lcl[i], n_lcl[i],         \
                                             q_rai[i], n_rai[i]);                                    \
            if (dq_lcl_dt) dq_lcl_dt[i] = o.dq_lcl_dt;                                               \
            if (dn_lcl_dt) dn_lcl_dt[i] = o.dn_lcl_dt;                                               \
            if (dq_rai_dt) dq_rai_dt[i] = o.dq_rai_dt;                                               \
            if (dn_rai_dt) dn_rai_dt[i] = o.dn_rai_dt;                                               \
            if (leaves)                                                                              \
                for (int k = 0; k < CUMICRO_SB2006_NLEAF; ++k)                                       \
                    if (leaves[k]) leaves[k][i] = o.leaf[k];                                         \
        }                                                                                            \
        return 0;                                                                                    \
    }
DEF_BMT2M_WARM(f64, double)
DEF_BMT2M_WARM(f32, float)

// Rounding-error bounds of the reference algorithm itself (oracle_tracked.hpp): same
// inputs, FT = Tr; returns only the bounds (the values equal the f64 run bit for bit).
int oracle_bmt2m_warm_bound_f64(const cumicro_params_2m_warm_f64* p, int64_t n, const double* rho, const double* T,
                                const double* q_tot, const double* q_lcl, const double* n_lcl, const double* q_rai,
                                const double* n_rai, double* const* bound4, double* const* leaf_bounds) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        Warm2MOut<Tr> o = bmt2m_warm<Tr>(*p, Tr(rho[i]), Tr(T[i]), Tr(q_tot[i]), Tr(q_lcl[i]), Tr(n_lcl[i]),
                                         Tr(q_rai[i]), Tr(n_rai[i]));
        if (bound4) {
            bound4[0][i] = o.dq_lcl_dt.e;
            bound4[1][i] = o.dn_lcl_dt.e;
            bound4[2][i] = o.dq_rai_dt.e;
            bound4[3][i] = o.dn_rai_dt.e;
        }
        if (leaf_bounds)
            for (int k = 0; k < CUMICRO_SB2006_NLEAF; ++k)
                if (leaf_bounds[k]) leaf_bounds[k][i] = o.leaf[k].e;
    }
    return 0;
}
#define DEF_TERMVEL_BOUND(NAME, PDF, VEL, FN)                                                                    \
    int oracle_##NAME##_bound_f64(const PDF* pdf, const VEL* vel, int64_t n, const double* q, const double* rho, \
                                  const double* N, double* b0, double* b1) {                                     \
        _Pragma("omp parallel for schedule(static)") for (int64_t i = 0; i < n; ++i) {                           \
            Tr v0, v1;                                                                                           \
            FN<Tr>(*pdf, *vel, Tr(q[i]), Tr(rho[i]), Tr(N[i]), v0, v1);                                          \
            b0[i] = v0.e;                                                                                        \
            b1[i] = v1.e;                                                                                        \
        }                                                                                                        \
        return 0;                                                                                                \
    }
DEF_TERMVEL_BOUND(termvel_2m_rain_sb, cumicro_sb_pdf_r_f64, cumicro_vel_sb2006_f64, rain_terminal_velocity_sb)
DEF_TERMVEL_BOUND(termvel_2m_rain_chen, cumicro_sb_pdf_r_f64, cumicro_vel_chen_rain_f64, rain_terminal_velocity_chen)
DEF_TERMVEL_BOUND(termvel_2m_cloud, cumicro_sb_pdf_c_f64, cumicro_vel_stokes_f64, cloud_terminal_velocity)

#define DEF_TERMVEL_2M(SUF, FT)                                                                      \
    int oracle_termvel_2m_rain_sb_##SUF(const cumicro_sb_pdf_r_##SUF* pdf,                           \
                                        const cumicro_vel_sb2006_##SUF* vel, int64_t n,              \
                                        const FT* q, const FT* rho, const FT* N, FT* vt0, FT* vt1) { \
        _Pragma("omp parallel for schedule(static)") for (int64_t i = 0; i < n; ++i)                 \
            rain_terminal_velocity_sb<FT>(*pdf, *vel, q[i], rho[i], N[i], vt0[i], vt1[i]);           \
        return 0;                                                                                    \
    }                                                                                                \
    int oracle_termvel_2m_rain_chen_##SUF(const cumicro_sb_pdf_r_##SUF* pdf,                         \
                                          const cumicro_vel_chen_rain_##SUF* vel, int64_t n,         \
                                          const FT* q, const FT* rho, const FT* N, FT* vt0,          \
                                          FT* vt1) {                                                 \
        _Pragma("omp parallel for schedule(static)") for (int64_t i = 0; i < n; ++i)                 \
            rain_terminal_velocity_chen<FT>(*pdf, *vel, q[i], rho[i], N[i], vt0[i], vt1[i]);         \
        return 0;                                                                                    \
    }                                                                                                \
    int oracle_termvel_2m_cloud_##SUF(const cumicro_sb_pdf_c_##SUF* pdf,                             \
                                      const cumicro_vel_stokes_##SUF* vel, int64_t n, const FT* q,   \
                                      const FT* rho, const FT* N, FT* vt0, FT* vt1) {                \
        _Pragma("omp parallel for schedule(static)") for (int64_t i = 0; i < n; ++i)                 \
            cloud_terminal_velocity<FT>(*pdf, *vel, q[i], rho[i], N[i], vt0[i], vt1[i]);             \
        return 0;                                                                                    \
    }
DEF_TERMVEL_2M(f64, double)
DEF_TERMVEL_2M(f32, float)

}  // extern "C"
