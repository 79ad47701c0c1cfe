// Host-side harness of cm_sb2006_fast.cuh for tests/test_warm2m_fast_host.py: evaluates the HOST instantiation of the headline
// kernel body (same arithmetic; MUFU seeds emulated) so its parity against the oracle is checked without a GPU.
#include "../../cloudmicrophysics.jl_b200/csrc/cm_sb2006_fast.cuh"

static int g_use_table = 1;
extern "C" {
void cmt_use_table(int on) { g_use_table = on; }
double cmt_table_error(const cumicro_params_2m_warm_f64* p) {
    cm::W2K k = cm::make_w2k(*p, false);
    static double tab[cm::kTabDoubles];
    return cm::build_w2_table(*p, k, tab);
}
int cmt_warm2m_supported(const cumicro_params_2m_warm_f64* p) { return cm::w2k_supported(*p) ? 1 : 0; }
void cmt_warm2m_fast(const cumicro_params_2m_warm_f64* p, int f32_method, long n, const double* rho, const double* T,
                     const double* q_tot, const double* q_lcl, const double* n_lcl, const double* q_rai, const double* n_rai,
                     double* o0, double* o1, double* o2, double* o3) {
    cm::W2K k = cm::make_w2k(*p, f32_method != 0);
    static double tab[cm::kTabDoubles];
    const bool use_tab = p->sb.pdf_r.limited && g_use_table && cm::build_w2_table(*p, k, tab) < 1e-15;
    for (long i = 0; i < n; ++i) {
        double y[4];
        if (use_tab)
            cm::warm2m_fast<1, true>(k, rho[i], T[i], q_tot[i], q_lcl[i], n_lcl[i], q_rai[i], n_rai[i], 0.0, false, y, tab);
        else if (p->sb.pdf_r.limited)
            cm::warm2m_fast<1>(k, rho[i], T[i], q_tot[i], q_lcl[i], n_lcl[i], q_rai[i], n_rai[i], 0.0, false, y);
        else
            cm::warm2m_fast<0>(k, rho[i], T[i], q_tot[i], q_lcl[i], n_lcl[i], q_rai[i], n_rai[i], 0.0, false, y);
        o0[i] = y[0]; o1[i] = y[1]; o2[i] = y[2]; o3[i] = y[3];
    }
}
void cmt_log_abs(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = cm::log_abs_(x[i]); }
}
