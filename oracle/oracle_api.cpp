// oracle_api.cpp — extern "C" array drivers over the scalar CPU restatement.
// TEST INFRASTRUCTURE ONLY (see oracle_base.hpp).  OpenMP over points: this is
// also the "port" CPU baseline timed by bench.py (the reference has no
// multithreaded driver of its own; hosts broadcast the scalar methods).
#include <omp.h>

#include "oracle_icenuc.hpp"

using namespace orc;

extern "C" {

int oracle_num_threads(void) { return omp_get_max_threads(); }
void oracle_set_num_threads(int n) { omp_set_num_threads(n); }
// 1: Float64 / error-bound evaluations use the Float32 method's thresholds (oracle_base.hpp)
void oracle_set_f32_thresholds(int on) { f32_thresholds() = (on != 0); }

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

// ---- 1-moment scheme ------------------------------------------------------------------------
// mode: 0 Instantaneous (out4), 1 InstantaneousVerbose (out4 + src18), 2 LinearizedAverage (out4; dt, nsub)
#define DEF_BMT1M(SUF, FT)                                                                                         \
    int oracle_bmt1m_##SUF(const cumicro_params_1m_##SUF* p, int mode, int64_t n, const FT* rho, const FT* T,      \
                           const FT* q_tot, const FT* q_lcl, const FT* q_icl, const FT* q_rai, const FT* q_sno,    \
                           FT dt, int nsub, FT* const* out4, FT* const* src18) {                                   \
        _Pragma("omp parallel for schedule(static)") for (int64_t i = 0; i < n; ++i) {                             \
            FT o[4];                                                                                               \
            if (mode == 2) {                                                                                       \
                bmt1m_linearized_average<FT>(*p, rho[i], T[i], q_tot[i], q_lcl[i], q_icl[i], q_rai[i], q_sno[i],   \
                                             dt, nsub, o);                                                         \
            } else {                                                                                               \
                Src1M<FT> r = microphysics_source_terms_1m<FT>(*p, rho[i], T[i], q_tot[i], q_lcl[i], q_icl[i],     \
                                                               q_rai[i], q_sno[i]);                                \
                aggregate_tendencies_1m<FT>(r, o);                                                                 \
                if (src18)                                                                                         \
                    for (int k = 0; k < S1M_NSRC; ++k)                                                             \
                        if (src18[k]) src18[k][i] = r.s[k];                                                        \
            }                                                                                                      \
            for (int k = 0; k < 4; ++k)                                                                            \
                if (out4 && out4[k]) out4[k][i] = o[k];                                                            \
        }                                                                                                          \
        return 0;                                                                                                  \
    }
DEF_BMT1M(f64, double)
DEF_BMT1M(f32, float)

int oracle_bmt1m_bound_f64(const cumicro_params_1m_f64* p, int mode, int64_t n, const double* rho, const double* T,
                           const double* q_tot, const double* q_lcl, const double* q_icl, const double* q_rai,
                           const double* q_sno, double dt, int nsub, double* const* out4, double* const* src18) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        Tr o[4];
        if (mode == 2) {
            bmt1m_linearized_average<Tr>(*p, Tr(rho[i]), Tr(T[i]), Tr(q_tot[i]), Tr(q_lcl[i]), Tr(q_icl[i]), Tr(q_rai[i]),
                                         Tr(q_sno[i]), Tr(dt), nsub, o);
        } else {
            Src1M<Tr> r = microphysics_source_terms_1m<Tr>(*p, Tr(rho[i]), Tr(T[i]), Tr(q_tot[i]), Tr(q_lcl[i]), Tr(q_icl[i]),
                                                           Tr(q_rai[i]), Tr(q_sno[i]));
            aggregate_tendencies_1m<Tr>(r, o);
            if (src18)
                for (int k = 0; k < S1M_NSRC; ++k)
                    if (src18[k]) src18[k][i] = r.s[k].e;
        }
        for (int k = 0; k < 4; ++k)
            if (out4 && out4[k]) out4[k][i] = o[k].e;
    }
    return 0;
}

// terminal velocities of the 1-moment / non-equilibrium schemes; kind:
// 0 rain Blk1M, 1 snow Blk1M, 2 rain Chen2022, 3 snow Chen2022 (large ice), 4 cloud liquid Stokes, 5 cloud ice Chen2022 (small ice)
#define DEF_TERMVEL_1M(SUF, FT)                                                                                    \
    int oracle_termvel_1m_##SUF(const cumicro_params_1m_##SUF* p, const void* vel, int kind, int64_t n,            \
                                const FT* rho, const FT* q, FT* out) {                                             \
        _Pragma("omp parallel for schedule(static)") for (int64_t i = 0; i < n; ++i) {                             \
            switch (kind) {                                                                                        \
                case 0: out[i] = terminal_velocity_1m_rain_blk<FT>(*p, rho[i], q[i]); break;                       \
                case 1: out[i] = terminal_velocity_1m_snow_blk<FT>(*p, rho[i], q[i]); break;                       \
                case 2: out[i] = terminal_velocity_1m_rain_chen<FT>(*p, *(const cumicro_vel_chen_rain_##SUF*)vel, rho[i], q[i]); break; \
                case 3: out[i] = terminal_velocity_1m_snow_chen<FT>(*p, *(const cumicro_vel_chen_large_ice_##SUF*)vel, rho[i], q[i]); break; \
                case 4: out[i] = terminal_velocity_noneq_liquid<FT>(*p, *(const cumicro_vel_stokes_##SUF*)vel, rho[i], q[i]); break; \
                case 5: out[i] = terminal_velocity_noneq_ice<FT>(*p, *(const cumicro_vel_chen_small_ice_##SUF*)vel, rho[i], q[i]); break; \
                default: out[i] = 0;                                                                               \
            }                                                                                                      \
        }                                                                                                          \
        return 0;                                                                                                  \
    }
DEF_TERMVEL_1M(f64, double)
DEF_TERMVEL_1M(f32, float)

int oracle_termvel_1m_bound_f64(const cumicro_params_1m_f64* p, const void* vel, int kind, int64_t n, const double* rho,
                                const double* q, double* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        Tr r(rho[i]), qq(q[i]), v;
        switch (kind) {
            case 0: v = terminal_velocity_1m_rain_blk<Tr>(*p, r, qq); break;
            case 1: v = terminal_velocity_1m_snow_blk<Tr>(*p, r, qq); break;
            case 2: v = terminal_velocity_1m_rain_chen<Tr>(*p, *(const cumicro_vel_chen_rain_f64*)vel, r, qq); break;
            case 3: v = terminal_velocity_1m_snow_chen<Tr>(*p, *(const cumicro_vel_chen_large_ice_f64*)vel, r, qq); break;
            case 4: v = terminal_velocity_noneq_liquid<Tr>(*p, *(const cumicro_vel_stokes_f64*)vel, r, qq); break;
            case 5: v = terminal_velocity_noneq_ice<Tr>(*p, *(const cumicro_vel_chen_small_ice_f64*)vel, r, qq); break;
            default: break;
        }
        out[i] = v.e;
    }
    return 0;
}

// ---- ice nucleation, water activity, ARG2000 ---------------------------------------------------
// what: 0 deposition_J, 1 ABIFM_J, 2 homogeneous_J_cubic, 3 homogeneous_J_linear (x = Δa_w)
//       4 a_w_ice(T), 5 a_w_eT(e = y, T = x), 6 a_w_xT(x_frac = y, T = x), 7 H2SO4 p_sol(x_frac = y, T = x)
//       8 P3_deposition_N_i(T), 9 INP_concentration_mean(T), 10 dust_activated_number_fraction(Si = x, T = y)
// returns the number of per-point domain errors (DomainError / AssertionError of the reference)
#define DEF_ICENUC(SUF, FT)                                                                                          \
    int64_t oracle_icenuc_##SUF(const cumicro_params_icenuc_##SUF* p, int what, int64_t n, const FT* x, const FT* y,    \
                                FT* out) {                                                                             \
        int64_t nerr = 0;                                                                                              \
        Thermo<FT> tps(p->tps);                                                                                        \
        _Pragma("omp parallel for schedule(static) reduction(+ : nerr)") for (int64_t i = 0; i < n; ++i) {             \
            bool err = false;                                                                                          \
            FT v = 0;                                                                                                  \
            switch (what) {                                                                                            \
                case 0: v = deposition_J<FT>(p->dust, x[i]); break;                                                    \
                case 1: v = ABIFM_J<FT>(p->dust, x[i]); break;                                                         \
                case 2: v = homogeneous_J_cubic<FT>(p->koop, x[i], err); break;                                        \
                case 3: v = homogeneous_J_linear<FT>(p->koop, x[i]); break;                                            \
                case 4: v = a_w_ice<FT>(tps, x[i]); break;                                                             \
                case 5: v = a_w_eT<FT>(tps, y[i], x[i]); break;                                                        \
                case 6: v = a_w_xT<FT>(p->h2so4, tps, y[i], x[i]); break;                                              \
                case 7: v = H2SO4_soln_saturation_vapor_pressure<FT>(p->h2so4, y[i], x[i]); break;                     \
                case 8: v = P3_deposition_N_i<FT>(p->mm2014, x[i]); break;                                             \
                case 9: v = INP_concentration_mean<FT>(p->frostenberg, x[i]); break;                                   \
                case 10: v = dust_activated_number_fraction<FT>(p->dust, p->mohler, x[i], y[i], err); break;           \
                default: break;                                                                                        \
            }                                                                                                          \
            if (err) { v = std::numeric_limits<FT>::quiet_NaN(); nerr += 1; }                                          \
            out[i] = v;                                                                                                \
        }                                                                                                              \
        return nerr;                                                                                                   \
    }                                                                                                                  \
    /* ARG2000 + (optionally) the nucleation rates at Δa_w = a_w_eT(p_v, T) - a_w_ice(T): columns T,p,w,q_tot,q_liq,    \
       q_ice,N_liq,N_ice -> S_max, N_act[n_modes], M_act[n_modes], J_dep, J_ABIFM, J_hom (any pointer may be NULL) */   \
    int64_t oracle_arg_icenuc_##SUF(const cumicro_params_icenuc_##SUF* p, int64_t n, const FT* T, const FT* pr,         \
                                    const FT* w, const FT* q_tot, const FT* q_liq, const FT* q_ice, const FT* N_liq,    \
                                    const FT* N_ice, FT* S_max, FT* const* N_act, FT* const* M_act, FT* J_dep,         \
                                    FT* J_abifm, FT* J_hom, FT* da_w_out) {                                            \
        int64_t nerr = 0;                                                                                              \
        _Pragma("omp parallel for schedule(static) reduction(+ : nerr)") for (int64_t i = 0; i < n; ++i) {             \
            FT na[8], ma[8], sm;                                                                                       \
            activated_per_mode<FT>(*p, T[i], pr[i], w[i], q_tot[i], q_liq[i], q_ice[i], N_liq[i], N_ice[i], sm, na,    \
                                   ma);                                                                                \
            if (S_max) S_max[i] = sm;                                                                                  \
            for (int k = 0; k < p->n_modes; ++k) {                                                                     \
                if (N_act && N_act[k]) N_act[k][i] = na[k];                                                            \
                if (M_act && M_act[k]) M_act[k][i] = ma[k];                                                            \
            }                                                                                                          \
            if (J_dep || J_abifm || J_hom || da_w_out) {                                                               \
                Thermo<FT> tps(p->tps);                                                                                \
                FT R_m = tps.R_m(q_tot[i], q_liq[i], q_ice[i]);                                                        \
                FT rho = pr[i] / (R_m * T[i]);                                                                         \
                FT e = (q_tot[i] - q_liq[i] - q_ice[i]) * rho * tps.R_v() * T[i];                                      \
                FT d = a_w_eT<FT>(tps, e, T[i]) - a_w_ice<FT>(tps, T[i]);                                              \
                if (da_w_out) da_w_out[i] = d;                                                                         \
                if (J_dep) J_dep[i] = deposition_J<FT>(p->dust, d);                                                    \
                if (J_abifm) J_abifm[i] = ABIFM_J<FT>(p->dust, d);                                                     \
                if (J_hom) {                                                                                           \
                    bool err = false;                                                                                  \
                    FT v = p->hom_linear ? homogeneous_J_linear<FT>(p->koop, d) : homogeneous_J_cubic<FT>(p->koop, d, err); \
                    if (err) { v = std::numeric_limits<FT>::quiet_NaN(); nerr += 1; }                                  \
                    J_hom[i] = v;                                                                                      \
                }                                                                                                      \
            }                                                                                                          \
        }                                                                                                              \
        return nerr;                                                                                                   \
    }
DEF_ICENUC(f64, double)
DEF_ICENUC(f32, float)

// rounding-error bounds of the reference algorithm for oracle_arg_icenuc_f64's outputs
int64_t oracle_arg_icenuc_bound_f64(const cumicro_params_icenuc_f64* p, int64_t n, const double* T, const double* pr,
                                    const double* w, const double* q_tot, const double* q_liq, const double* q_ice,
                                    const double* N_liq, const double* N_ice, double* S_max, double* const* N_act,
                                    double* const* M_act, double* J_dep, double* J_abifm, double* J_hom, double* da_w_out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        Tr na[8], ma[8], sm;
        activated_per_mode<Tr>(*p, Tr(T[i]), Tr(pr[i]), Tr(w[i]), Tr(q_tot[i]), Tr(q_liq[i]), Tr(q_ice[i]), Tr(N_liq[i]), Tr(N_ice[i]),
                               sm, na, ma);
        if (S_max) S_max[i] = sm.e;
        for (int k = 0; k < p->n_modes; ++k) {
            if (N_act && N_act[k]) N_act[k][i] = na[k].e;
            if (M_act && M_act[k]) M_act[k][i] = ma[k].e;
        }
        if (J_dep || J_abifm || J_hom || da_w_out) {
            Thermo<Tr> tps(p->tps);
            Tr R_m = tps.R_m(Tr(q_tot[i]), Tr(q_liq[i]), Tr(q_ice[i]));
            Tr rho = Tr(pr[i]) / (R_m * Tr(T[i]));
            Tr e = (Tr(q_tot[i]) - Tr(q_liq[i]) - Tr(q_ice[i])) * rho * tps.R_v() * Tr(T[i]);
            Tr d = a_w_eT<Tr>(tps, e, Tr(T[i])) - a_w_ice<Tr>(tps, Tr(T[i]));
            if (da_w_out) da_w_out[i] = d.e;
            if (J_dep) J_dep[i] = deposition_J<Tr>(p->dust, d).e;
            if (J_abifm) J_abifm[i] = ABIFM_J<Tr>(p->dust, d).e;
            if (J_hom) {
                bool err = false;
                J_hom[i] = p->hom_linear ? homogeneous_J_linear<Tr>(p->koop, d).e : homogeneous_J_cubic<Tr>(p->koop, d, err).e;
            }
        }
    }
    return 0;
}

}  // extern "C"
