/* Plain-C caller of libcumicro.so: what a non-Julia, non-Python host does.  Fills cumicro_params_2m_warm_f64 BY HAND (the default
 * SB2006 block of the reference), prints the offsetof / sizeof table that tests/test_abi.py compares with ctypes,
 * and — when a CUDA device is present (argv[1] = "run") — calls cumicro_bmt2m_warm_f64 on the golden state of
 * test/gpu_tests.jl:844-871 through the CUDA runtime it loads with dlopen (no CUDA headers needed to build this file).
 *   gcc -I include tests/native/c_harness.c -ldl -o c_harness && ./c_harness offsets */
#include <dlfcn.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cumicro.h"

#define OFF(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))

static void print_offsets(void) {
    printf("sizeof cumicro_params_2m_warm_f64 %zu\n", sizeof(cumicro_params_2m_warm_f64));
    OFF(cumicro_params_2m_warm_f64, tps); OFF(cumicro_params_2m_warm_f64, sb); OFF(cumicro_params_2m_warm_f64, aps);
    OFF(cumicro_params_2m_warm_f64, condevap_tau_relax); OFF(cumicro_params_2m_warm_f64, subdep_tau_relax);
    OFF(cumicro_sb2006_f64, pdf_c); OFF(cumicro_sb2006_f64, pdf_r); OFF(cumicro_sb2006_f64, acnv); OFF(cumicro_sb2006_f64, accr);
    OFF(cumicro_sb2006_f64, self); OFF(cumicro_sb2006_f64, brek); OFF(cumicro_sb2006_f64, evap); OFF(cumicro_sb2006_f64, numadj_tau);
    OFF(cumicro_sb_pdf_r_f64, rho0); OFF(cumicro_sb_pdf_r_f64, limited); OFF(cumicro_sb_pdf_r_f32, rho0); OFF(cumicro_sb_pdf_r_f32, limited);
    OFF(cumicro_thermo_f64, grav); OFF(cumicro_thermo_f32, grav);
    OFF(cumicro_params_1m_f64, cloud_ice); OFF(cumicro_params_1m_f64, snow); OFF(cumicro_params_1m_f64, pp); OFF(cumicro_params_1m_f64, processes);
    OFF(cumicro_params_1m_f32, pp); OFF(cumicro_params_1m_f32, processes);
    OFF(cumicro_params_icenuc_f64, modes); OFF(cumicro_params_icenuc_f64, n_modes); OFF(cumicro_params_icenuc_f32, n_modes);
    OFF(cumicro_params_p3_f64, scheme); OFF(cumicro_params_p3_f64, quad); OFF(cumicro_quadrature_f64, n); OFF(cumicro_quadrature_f32, n);
    OFF(cumicro_p3_scheme_f64, slope_power_law); OFF(cumicro_p3_scheme_f32, slope_power_law);
}

static void fill_defaults(cumicro_params_2m_warm_f64* p) {
    memset(p, 0, sizeof(*p));
    /* ThermodynamicsParameters (ClimaParams 1.0.18 defaults) */
    p->tps.T_0 = 273.16; p->tps.T_triple = 273.16; p->tps.press_triple = 611.657; p->tps.T_freeze = 273.15;
    p->tps.R_v = 461.5; p->tps.R_d = 287.0; p->tps.cp_d = 1004.5; p->tps.cp_v = 1859.0; p->tps.cp_l = 4181.0; p->tps.cp_i = 2070.0;
    p->tps.LH_v0 = 2500800.0; p->tps.LH_s0 = 2834400.0; p->tps.q_min = 1e-10; p->tps.grav = 9.81;
    /* AirProperties */
    p->aps.K_therm = 0.024; p->aps.D_vapor = 2.26e-5; p->aps.nu_air = 1.6e-5;
    /* SB2006 (src/parameters/Microphysics2M.jl:314-672) */
    cumicro_sb2006_f64* sb = &p->sb;
    sb->pdf_c.nu_c = 1.0; sb->pdf_c.mu_c = 1.0; sb->pdf_c.xc_min = 4.2e-15; sb->pdf_c.xc_max = 2.6e-10; sb->pdf_c.rho_w = 1000.0;
    sb->pdf_c.loggamma_z1 = 0.0;                 /* loggamma((nu+1)/mu) = loggamma(2) */
    sb->pdf_c.loggamma_z2 = 0.693147180559945;   /* loggamma((nu+2)/mu) = loggamma(3) */
    sb->pdf_r.nu_r = -2.0 / 3.0; sb->pdf_r.mu_r = 1.0 / 3.0; sb->pdf_r.xr_min = 2.6e-10; sb->pdf_r.xr_max = 5e-6;
    sb->pdf_r.N0_min = 2.5e5; sb->pdf_r.N0_max = 2e7; sb->pdf_r.lam_min = 1e3; sb->pdf_r.lam_max = 1e4;
    sb->pdf_r.rho_w = 1000.0; sb->pdf_r.rho0 = 1.225; sb->pdf_r.limited = 1;
    sb->acnv.kcc = 4.44e9; sb->acnv.x_star = 2.6e-10; sb->acnv.rho0 = 1.225; sb->acnv.A = 400.0; sb->acnv.a = 0.7; sb->acnv.b = 3.0;
    sb->accr.kcr = 5.25; sb->accr.tau0 = 5e-5; sb->accr.rho0 = 1.225; sb->accr.c = 4.0;
    sb->self.krr = 7.12; sb->self.kappa_rr = 60.7; sb->self.d = -5.0;
    sb->brek.Deq = 9e-4; sb->brek.Dr_th = 3.5e-4; sb->brek.kbr = 1000.0; sb->brek.kappa_br = 2300.0;
    /* EvaporationSB2006 with the five host-side ventilation constants of Microphysics2M.jl:566-575 */
    sb->evap.av = 0.78; sb->evap.bv = 0.308; sb->evap.alpha = 159.0; sb->evap.beta = 0.266; sb->evap.rho0 = 1.225;
    sb->evap.a_vent_1 = 0.4292505423563015; sb->evap.b_vent_1 = 0.18089257644312223;
    sb->evap.a_vent_0_coeff = 2.575503254137809; sb->evap.b_vent_0_coeff = 0.5944731244808867; sb->evap.beta_vent_0 = -0.10099999999999998;
    sb->numadj_tau = 100.0;
    p->condevap_tau_relax = 10.0; p->subdep_tau_relax = 10.0;
}

typedef int (*cuda_malloc_t)(void**, size_t);
typedef int (*cuda_memcpy_t)(void*, const void*, size_t, int);
typedef int (*cuda_sync_t)(void);

int main(int argc, char** argv) {
    if (argc < 2 || strcmp(argv[1], "offsets") == 0) { print_offsets(); return 0; }
    /* "run <libcumicro.so> <params.bin>": the parameter block comes from the file when given (the Python test writes the packed
     * default block), otherwise from fill_defaults() above */
    const char* libpath = argc > 2 ? argv[2] : "libcumicro.so";
    void* h = dlopen(libpath, RTLD_NOW);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    int (*bmt)(const cumicro_params_2m_warm_f64*, int64_t, const double*, const double*, const double*, const double*, const double*,
               const double*, const double*, const double*, double*, double*, double*, double*, double* const*, void*) =
        (int (*)(const cumicro_params_2m_warm_f64*, int64_t, const double*, const double*, const double*, const double*, const double*,
                 const double*, const double*, const double*, double*, double*, double*, double*, double* const*, void*))dlsym(h, "cumicro_bmt2m_warm_f64");
    const char* (*last_error)(void) = (const char* (*)(void))dlsym(h, "cumicro_last_error");
    void* rt = dlopen("libcudart.so.12", RTLD_NOW);
    if (!rt) rt = dlopen("libcudart.so", RTLD_NOW);
    if (!bmt || !rt) { fprintf(stderr, "symbols / CUDA runtime not found\n"); return 2; }
    cuda_malloc_t cuda_malloc = (cuda_malloc_t)dlsym(rt, "cudaMalloc");
    cuda_memcpy_t cuda_memcpy = (cuda_memcpy_t)dlsym(rt, "cudaMemcpy");
    cuda_sync_t cuda_sync = (cuda_sync_t)dlsym(rt, "cudaDeviceSynchronize");
    cumicro_params_2m_warm_f64 p;
    fill_defaults(&p);
    if (argc > 3) {
        FILE* f = fopen(argv[3], "rb");
        if (!f || fread(&p, sizeof(p), 1, f) != 1) { fprintf(stderr, "cannot read %s\n", argv[3]); return 2; }
        fclose(f);
    }
    /* golden state of test/gpu_tests.jl:821-843: T = 290, q_tot = 7e-3, q_lcl = 2e-3, q_rai = 5e-4, rho = 1.2, N_lcl = 1e8, N_rai = 1e7 */
    enum { N = 4 };
    const double rho = 1.2;
    const double host_in[7] = {rho, 290.0, 7e-3, 2e-3, 1e8 / rho, 5e-4, 1e7 / rho};
    double* d_in[7];
    double* d_out[4];
    for (int c = 0; c < 7; ++c) {
        double col[N];
        for (int i = 0; i < N; ++i) col[i] = host_in[c];
        if (cuda_malloc((void**)&d_in[c], sizeof(col)) != 0) { fprintf(stderr, "cudaMalloc failed (no device?)\n"); return 3; }
        cuda_memcpy(d_in[c], col, sizeof(col), 1 /* cudaMemcpyHostToDevice */);
    }
    for (int c = 0; c < 4; ++c) cuda_malloc((void**)&d_out[c], sizeof(double) * N);
    const int st = bmt(&p, N, d_in[0], d_in[1], d_in[2], d_in[3], d_in[4], d_in[5], d_in[6], NULL, d_out[0], d_out[1], d_out[2], d_out[3],
                       NULL, NULL);
    if (st != 0) { fprintf(stderr, "cumicro_bmt2m_warm_f64 -> %d: %s\n", st, last_error ? last_error() : "?"); return 4; }
    cuda_sync();
    for (int c = 0; c < 4; ++c) {
        double col[N];
        cuda_memcpy(col, d_out[c], sizeof(col), 2 /* cudaMemcpyDeviceToHost */);
        printf("out%d %.17g\n", c, col[0]);
    }
    return 0;
}
