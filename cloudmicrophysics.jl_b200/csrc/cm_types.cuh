// cm_types.cuh — float-type -> parameter-block type map, launch helpers.
#pragma once
#include "../../include/cumicro.h"
#include "cm_math.cuh"
#include "cm_widen.inc"

namespace cm {

template <class FT> struct P;
template <> struct P<double> {
    using thermo = cumicro_thermo_f64;
    using air = cumicro_air_f64;
    using sb_pdf_c = cumicro_sb_pdf_c_f64;
    using sb_pdf_r = cumicro_sb_pdf_r_f64;
    using sb2006 = cumicro_sb2006_f64;
    using vel_sb2006 = cumicro_vel_sb2006_f64;
    using vel_stokes = cumicro_vel_stokes_f64;
    using vel_chen_rain = cumicro_vel_chen_rain_f64;
    using vel_chen_small_ice = cumicro_vel_chen_small_ice_f64;
    using vel_chen_large_ice = cumicro_vel_chen_large_ice_f64;
    using params_2m_warm = cumicro_params_2m_warm_f64;
    using params_1m = cumicro_params_1m_f64;
    using particle_mass = cumicro_particle_mass_f64;
    using frostenberg = cumicro_frostenberg2023_f64;
};
template <> struct P<float> {
    using thermo = cumicro_thermo_f32;
    using air = cumicro_air_f32;
    using sb_pdf_c = cumicro_sb_pdf_c_f32;
    using sb_pdf_r = cumicro_sb_pdf_r_f32;
    using sb2006 = cumicro_sb2006_f32;
    using vel_sb2006 = cumicro_vel_sb2006_f32;
    using vel_stokes = cumicro_vel_stokes_f32;
    using vel_chen_rain = cumicro_vel_chen_rain_f32;
    using vel_chen_small_ice = cumicro_vel_chen_small_ice_f32;
    using vel_chen_large_ice = cumicro_vel_chen_large_ice_f32;
    using params_2m_warm = cumicro_params_2m_warm_f32;
    using params_1m = cumicro_params_1m_f32;
    using particle_mass = cumicro_particle_mass_f32;
    using frostenberg = cumicro_frostenberg2023_f32;
};

}  // namespace cm

// ---- host-side plumbing shared by the .cu files (defined in cm_host.cu) ----------
namespace cmh {
int fail(int code, const char* fmt, ...);  // records cumicro_last_error, returns code
int cuda_status(cudaError_t e, const char* what);
void count_launch(int n = 1);
bool aligned16(const void* p);
int num_sms();
}  // namespace cmh
