// oracle_api.cpp — extern "C" array drivers over the scalar CPU restatement.
// TEST INFRASTRUCTURE ONLY (see oracle_base.hpp).  OpenMP over points: this is
// also the "port" CPU baseline timed by bench.py (the reference has no
// multithreaded driver of its own; hosts broadcast the scalar methods).
#include <omp.h>

#include "oracle_p3.hpp"
#include "oracle_diag.hpp"

using namespace orc;

// CM2.rain_evaporation + CM2.∂rain_evaporation_∂N_rai_∂q_rai (CM2:780-853): in8 = q_tot, q_lcl, q_icl, q_rai, q_sno, rho, N_rai, T
template <class FT, class PB>
static void rain_evap_point(const PB& p, const FT (&x)[8], FT (&y)[4]) {
    Thermo<FT> tps(p.tps);
    FT dn, dq;
    rain_evaporation<FT>(p.sb, p.aps, tps, x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], dn, dq);
    y[0] = dn;
    y[1] = dq;
    y[2] = (x[6] > eps_2M<FT>()) ? dn / x[6] : FT(0);     // CM2:850
    y[3] = (x[3] > eps_2M<FT>()) ? dq / x[3] : FT(0);     // CM2:851
}
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

#define DEF_RAIN_EVAP(SUF, FT)                                                                                              \
    int oracle_rain_evaporation_2m_##SUF(const cumicro_params_2m_warm_##SUF* p, int64_t n, const FT* const* in8, FT* const* out4) { \
        _Pragma("omp parallel for schedule(static)") for (int64_t i = 0; i < n; ++i) {                                    \
            FT x[8], y[4];                                                                                                  \
            for (int c = 0; c < 8; ++c) x[c] = in8[c][i];                                                                   \
            rain_evap_point<FT>(*p, x, y);                                                                                  \
            for (int c = 0; c < 4; ++c) if (out4[c]) out4[c][i] = y[c];                                                     \
        }                                                                                                                   \
        return 0;                                                                                                           \
    }
DEF_RAIN_EVAP(f64, double)
DEF_RAIN_EVAP(f32, float)

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

// ---- P3 (Float64 and error-bound evaluations; parameters are the Float64 block) -----------------
// leaf: what = 0 gamma_inc P(a = x, x = y)   1 gamma_inc_inv(a = x, p = y, q = 1 - y)
//              2 rime_mass_fraction(q_rim = x, q_ice = y)   3 rime_density(q_rim = x, b_rim = y)
}  // extern "C"
template <class FT> static void p3_leaf(int what, int64_t n, const double* x, const double* y, double* out, bool bound) {
    for (int64_t i = 0; i < n; ++i) {
        FT r = FT(0);
        if (what == 0) { FT P, Q; gamma_inc<FT>(FT(x[i]), FT(y[i]), P, Q); r = P; }
        else if (what == 1) r = gamma_inc_inv<FT>(FT(x[i]), FT(y[i]), FT(1) - FT(y[i]));
        else if (what == 2) r = regularised_ratio<FT>(jmin(FT(x[i]), FT(y[i])), FT(y[i]));
        else if (what == 3) r = regularised_ratio<FT>(FT(x[i]), FT(y[i]));
        out[i] = bound ? Tr(r).e : val_(r);
    }
}
extern "C" {
int oracle_p3_leaf_f64(int what, int64_t n, const double* x, const double* y, double* out, int bound) {
    if (bound) p3_leaf<Tr>(what, n, x, y, out, true); else p3_leaf<double>(what, n, x, y, out, false);
    return 0;
}

// state-level functions over columns (L_ice, N_ice, F_rim | L_rim, rho_rim | B_rim, ...):
// from_prognostic = 1: the third / fourth columns are (L_rim, B_rim) and the state comes from state_from_prognostic
// outputs (each optional): thresholds[5] = rho_g, D_th, D_gr, D_cr, F_rim ; logl ; D_m ; v_n, v_m ; melt dN, dL ; self-collection dN ;
// collisions[10] ; sources[7] ; max_freeze(Dbar) ; rime_density_local(Dbar, Dbar)
}  // extern "C"
template <class FT>
static void p3_state_fns(const cumicro_params_p3_f64* p, int64_t n, int from_prognostic, const double* L_ice, const double* N_ice,
                         const double* c3, const double* c4, const double* rho_a, const double* T, const double* logl_in,
                         const double* L_c, const double* N_c, const double* L_r, const double* N_r, int logl_iters,
                         double* const* thr, double* logl_out, double* Dm, double* v_n, double* v_m, double* melt_dN,
                         double* melt_dL, double* selfcol, double* const* coll10, double* const* src7, double* maxfrz,
                         double* rimeloc, bool bound) {
    auto put = [&](double* dst, int64_t i, FT v) { if (dst) dst[i] = bound ? Tr(v).e : val_(v); };
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < n; ++i) {
        P3State<FT> s = from_prognostic ? state_from_prognostic<FT>(p->scheme, FT(L_ice[i]), FT(N_ice[i]), FT(c3[i]), FT(c4[i]))
                                        : make_p3_state<FT>(p->scheme, FT(L_ice[i]), FT(N_ice[i]), FT(c3[i]), FT(c4[i]));
        if (thr) { put(thr[0], i, s.rho_g); put(thr[1], i, s.D_th); put(thr[2], i, s.D_gr); put(thr[3], i, s.D_cr); put(thr[4], i, s.F_rim); }
        FT logl = logl_in ? FT(logl_in[i]) : get_distribution_loglambda<FT>(s, logl_iters);
        put(logl_out, i, logl);
        if (Dm) put(Dm, i, D_m<FT>(s, logl));
        FT ra = rho_a ? FT(rho_a[i]) : FT(1.2), Ta = T ? FT(T[i]) : FT(270);
        if (v_n || v_m) { FT a, b; ice_terminal_velocity_weighted<FT>(*p, ra, s, logl, FT(1e-6), a, b); put(v_n, i, a); put(v_m, i, b); }
        if (melt_dN || melt_dL) { FT a, b; ice_melt<FT>(*p, Ta, ra, s, logl, a, b); put(melt_dN, i, a); put(melt_dL, i, b); }
        if (selfcol) put(selfcol, i, ice_self_collection<FT>(*p, s, logl, ra));
        if (coll10) {
            Vec10<FT> r = liquid_ice_collisions<FT>(*p, s, logl, FT(L_c[i]), FT(N_c[i]), FT(L_r[i]), FT(N_r[i]), ra, Ta);
            for (int k = 0; k < 10; ++k) put(coll10[k], i, r.v[k]);
        }
        if (src7) {
            CollisionSources<FT> c = bulk_liquid_ice_collision_sources<FT>(*p, s, logl, FT(L_c[i]), FT(N_c[i]), FT(L_r[i]), FT(N_r[i]), ra, Ta);
            FT v[7] = {c.dq_c, c.dq_r, c.dN_c, c.dN_r, c.dL_rim, c.dL_ice, c.dB_rim};
            for (int k = 0; k < 7; ++k) put(src7[k], i, v[k]);
        }
        if (maxfrz || rimeloc) {
            using PP = cumicro_params_p3_f64;
            CollisionCtx<FT, PP> c;
            c.p = p; c.s = &s; c.rho_a = ra; c.T = Ta; c.rho_w = FT(p->warm.sb.pdf_c.rho_w);
            c.v_ice = ice_particle_terminal_velocity<FT>(*p, ra, s);
            c.v_liq = rain_particle_terminal_velocity<FT>(*p, ra);
            FT Dbar = exp_(-logl);
            put(maxfrz, i, max_freeze_rate<FT>(*p, c, Dbar));
            put(rimeloc, i, c.rho_rim_local(Dbar, Dbar));
        }
    }
}
extern "C" {
int oracle_p3_state_f64(const cumicro_params_p3_f64* p, int64_t n, int from_prognostic, const double* L_ice, const double* N_ice,
                        const double* c3, const double* c4, const double* rho_a, const double* T, const double* logl_in,
                        const double* L_c, const double* N_c, const double* L_r, const double* N_r, int logl_iters,
                        double* const* thr, double* logl_out, double* Dm, double* v_n, double* v_m, double* melt_dN, double* melt_dL,
                        double* selfcol, double* const* coll10, double* const* src7, double* maxfrz, double* rimeloc, int bound) {
    if (bound)
        p3_state_fns<Tr>(p, n, from_prognostic, L_ice, N_ice, c3, c4, rho_a, T, logl_in, L_c, N_c, L_r, N_r, logl_iters, thr, logl_out, Dm,
                         v_n, v_m, melt_dN, melt_dL, selfcol, coll10, src7, maxfrz, rimeloc, true);
    else
        p3_state_fns<double>(p, n, from_prognostic, L_ice, N_ice, c3, c4, rho_a, T, logl_in, L_c, N_c, L_r, N_r, logl_iters, thr, logl_out,
                             Dm, v_n, v_m, melt_dN, melt_dL, selfcol, coll10, src7, maxfrz, rimeloc, false);
    return 0;
}

// BMT:898-1083 over columns: in13 = rho,T,q_tot,q_lcl,n_lcl,q_rai,n_rai,q_ice,n_ice,q_rim,b_rim,logl,inpc_log_shift ; out9
}  // extern "C"
template <class FT> static void bmt2m_p3_cols(const cumicro_params_p3_f64* p, int64_t n, const double* const* in, double* const* out, bool bound) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < n; ++i) {
        BMT2MP3Out<FT> o = bmt2m_p3<FT>(*p, FT(in[0][i]), FT(in[1][i]), FT(in[2][i]), FT(in[3][i]), FT(in[4][i]), FT(in[5][i]), FT(in[6][i]),
                                        FT(in[7][i]), FT(in[8][i]), FT(in[9][i]), FT(in[10][i]), FT(in[11][i]), in[12] ? FT(in[12][i]) : FT(0));
        FT v[9] = {o.dq_lcl, o.dn_lcl, o.dq_rai, o.dn_rai, o.dq_ice, o.dn_ice, o.dq_rim, o.db_rim, o.dn_act};
        for (int k = 0; k < 9; ++k)
            if (out[k]) out[k][i] = bound ? Tr(v[k]).e : val_(v[k]);
    }
}
extern "C" {
int oracle_bmt2m_p3_f64(const cumicro_params_p3_f64* p, int64_t n, const double* const* in13, double* const* out9, int bound) {
    if (bound) bmt2m_p3_cols<Tr>(p, n, in13, out9, true); else bmt2m_p3_cols<double>(p, n, in13, out9, false);
    return 0;
}


// ---- multi-argument ice-nucleation rates: out[i] (and out2[i]) = fn(in[0][i], ...)
//   what 0: IN.MohlerDepositionRate(dust, mohler, Si, T, dSi_dt, N_aer)             IN:68-77
//        1: IN.P3_het_N_i(mm2014, T, N_l, V_l, dt)                                  IN:202-205
//        2: IN.INP_concentration_frequency(frostenberg, INPC, T)                    IN:219-224
//        3: P3.het_ice_nucleation(dust, tps, q_lcl, N_lcl, RH, T, rho) -> dNdt, dLdt   P3_processes.jl:20-45
}  // extern "C"
template <class FT, class PB>
static int64_t icenuc_rates_cols(const PB* p, int what, int64_t n, const FT* const* in, FT* out, FT* out2) {
    int64_t nerr = 0;
    Thermo<FT> tps(p->tps);
    for (int64_t i = 0; i < n; ++i) {
        bool err = false;
        FT v = FT(0), v2 = FT(0);
        switch (what) {
            case 0: v = MohlerDepositionRate<FT>(p->dust, p->mohler, in[0][i], in[1][i], in[2][i], in[3][i], err); break;
            case 1: v = P3_het_N_i<FT>(p->mm2014, in[0][i], in[1][i], in[2][i], in[3][i]); break;
            case 2: v = INP_concentration_frequency<FT>(p->frostenberg, in[0][i], in[1][i]); break;
            case 3: p3_het_ice_nucleation<FT>(p->dust, tps, in[0][i], in[1][i], in[2][i], in[3][i], in[4][i], v, v2); break;
            default: break;
        }
        if (err) { v = std::numeric_limits<FT>::quiet_NaN(); nerr += 1; }
        if (out) out[i] = v;
        if (out2) out2[i] = v2;
    }
    return nerr;
}
extern "C" {
int64_t oracle_icenuc_rates_f64(const cumicro_params_icenuc_f64* p, int what, int64_t n, const double* const* in, double* out, double* out2) {
    return icenuc_rates_cols<double>(p, what, n, in, out, out2);
}
int64_t oracle_icenuc_rates_f32(const cumicro_params_icenuc_f32* p, int what, int64_t n, const float* const* in, float* out, float* out2) {
    return icenuc_rates_cols<float>(p, what, n, in, out, out2);
}

// ---- F23 / Bigg nucleation rates as BMT:998-1075 calls them: in9 = rho,T,q_tot,q_lcl,n_lcl,q_rai,n_rai,q_ice,n_ice (+ shift)
}  // extern "C"
template <class FT> static void icenuc_f23_cols(const cumicro_params_p3_f64* p, int64_t n, const double* const* in, const double* shift, double* const* out, bool bound) {
    for (int64_t i = 0; i < n; ++i) {
        FT rho = clamp_to_nonneg(FT(in[0][i])), T = FT(in[1][i]), q_tot = clamp_to_nonneg(FT(in[2][i])), q_lcl = clamp_to_nonneg(FT(in[3][i])),
           n_lcl = clamp_to_nonneg(FT(in[4][i])), q_rai = clamp_to_nonneg(FT(in[5][i])), n_rai = clamp_to_nonneg(FT(in[6][i])),
           q_ice = clamp_to_nonneg(FT(in[7][i])), n_ice = clamp_to_nonneg(FT(in[8][i]));
        FT sh = shift ? FT(shift[i]) : FT(0);
        FT v[7];
        rain_freezing_rate<FT>(*p, q_rai, rho, FT(n_rai * rho), T, v[0], v[1]);
        cloud_freezing_rate<FT>(*p, q_lcl, rho, FT(n_lcl * rho), T, v[2], v[3]);
        v[4] = immersion_limit_rate<FT>(p->ice_nucleation, T, rho, FT(p->tau_act), sh, n_ice);
        FT m_nuc = FT(p->scheme.rho_i) * volume_sphere_D<FT>(FT(10e-6));
        f23_deposition_rate<FT>(*p, T, rho, q_tot, FT(q_lcl + q_rai), q_ice, n_ice, m_nuc, FT(p->tau_act), sh, v[5], v[6]);
        for (int k = 0; k < 7; ++k)
            if (out[k]) out[k][i] = bound ? Tr(v[k]).e : val_(v[k]);
    }
}
extern "C" {
int oracle_icenuc_f23_f64(const cumicro_params_p3_f64* p, int64_t n, const double* const* in9, const double* shift, double* const* out7, int bound) {
    if (bound) icenuc_f23_cols<Tr>(p, n, in9, shift, out7, true); else icenuc_f23_cols<double>(p, n, in9, shift, out7, false);
    return 0;
}

// ---- alternative 2-moment closures (KK2000, B1994, TC1980, LD2004)                       CM2:920-1002
}  // extern "C"
template <class FT, class PB> static void alt2m_cols(const PB* p, int what, int smooth, int64_t n, const FT* q_lcl, const FT* q_rai, const FT* rho, const FT* N_d, FT* out) {
    for (int64_t i = 0; i < n; ++i)
        out[i] = alt_2m<FT>(*p, what, smooth != 0, q_lcl[i], q_rai ? q_rai[i] : FT(0), rho ? rho[i] : FT(1), N_d ? N_d[i] : FT(1));
}
extern "C" {
int oracle_2m_alt_f64(const cumicro_params_2m_alt_f64* p, int what, int smooth, int64_t n, const double* q_lcl, const double* q_rai, const double* rho, const double* N_d, double* out) {
    alt2m_cols<double>(p, what, smooth, n, q_lcl, q_rai, rho, N_d, out);
    return 0;
}
int oracle_2m_alt_f32(const cumicro_params_2m_alt_f32* p, int what, int smooth, int64_t n, const float* q_lcl, const float* q_rai, const float* rho, const float* N_d, float* out) {
    alt2m_cols<float>(p, what, smooth, n, q_lcl, q_rai, rho, N_d, out);
    return 0;
}

// ---- cloud diagnostics (src/CloudDiagnostics.jl)
#define DEF_DIAG(SUF, FT)                                                                                                              \
    int oracle_diag_2m_##SUF(const cumicro_sb_pdf_c_##SUF* pc, const cumicro_sb_pdf_r_##SUF* pr, int64_t n, const FT* q_lcl,            \
                             const FT* q_rai, const FT* N_lcl, const FT* N_rai, const FT* rho, FT* Z, FT* reff) {                      \
        for (int64_t i = 0; i < n; ++i) {                                                                                              \
            if (Z) Z[i] = radar_reflectivity_2M<FT>(*pc, *pr, q_lcl[i], q_rai[i], N_lcl[i], N_rai[i], rho[i]);                         \
            if (reff) reff[i] = effective_radius_2M<FT>(*pc, *pr, q_lcl[i], q_rai[i], N_lcl[i], N_rai[i], rho[i]);                     \
        }                                                                                                                              \
        return 0;                                                                                                                      \
    }                                                                                                                                  \
    int oracle_diag_1m_##SUF(const cumicro_params_1m_##SUF* p, int64_t n, const FT* q_rai, const FT* rho, FT* Z) {                     \
        for (int64_t i = 0; i < n; ++i) Z[i] = radar_reflectivity_1M<FT>(*p, q_rai[i], rho[i]);                                        \
        return 0;                                                                                                                      \
    }                                                                                                                                  \
    int oracle_diag_reff_lh97_##SUF(FT rho_w, int64_t n, const FT* rho, const FT* q_lcl, const FT* N_lcl, const FT* q_rai,             \
                                    const FT* N_rai, FT* reff) {                                                                       \
        for (int64_t i = 0; i < n; ++i)                                                                                                \
            reff[i] = effective_radius_Liu_Hallet_97<FT>(rho_w, rho[i], q_lcl[i], N_lcl ? N_lcl[i] : FT(100), q_rai ? q_rai[i] : FT(0), \
                                                         N_rai ? N_rai[i] : FT(0));                                                    \
        return 0;                                                                                                                      \
    }
DEF_DIAG(f64, double)
DEF_DIAG(f32, float)

// ---- 0-moment scheme: BMT:658-680 -> CM0.remove_precipitation (src/Microphysics0M.jl:35-46), native FT arithmetic
}  // extern "C"
template <class FT, class PB> static void bmt0m_cols(const PB* p, int64_t n, const FT* q_lcl, const FT* q_icl, const FT* q_vap_sat, FT* out) {
    for (int64_t i = 0; i < n; ++i) {
        FT ql = clamp_to_nonneg(q_lcl[i]), qi = clamp_to_nonneg(q_icl[i]);
        FT thr = q_vap_sat ? FT(p->S_0 * q_vap_sat[i]) : FT(p->qc_0);
        out[i] = -jmax(FT(0), FT(ql + qi - thr)) / FT(p->tau_precip);
    }
}
extern "C" {
int oracle_bmt0m_f64(const cumicro_params_0m_f64* p, int64_t n, const double* q_lcl, const double* q_icl, const double* q_vap_sat, double* out) {
    bmt0m_cols<double>(p, n, q_lcl, q_icl, q_vap_sat, out);
    return 0;
}
int oracle_bmt0m_f32(const cumicro_params_0m_f32* p, int64_t n, const float* q_lcl, const float* q_icl, const float* q_vap_sat, float* out) {
    bmt0m_cols<float>(p, n, q_lcl, q_icl, q_vap_sat, out);
    return 0;
}

}  // extern "C"
