// oracle_diag.hpp — CPU restatement of src/CloudDiagnostics.jl (radar reflectivity, effective radius).
// Test infrastructure only (see oracle_base.hpp); operation order of the Julia source.
#pragma once
#include "oracle_1m.hpp"
#include "oracle_2m.hpp"

namespace orc {

// CMD.radar_reflectivity_1M((; pdf, mass)::Rain, q, ρ)                              CloudDiagnostics.jl:30-42
template <class FT, class P1> inline FT radar_reflectivity_1M(const P1& mp, FT q, FT rho) {
    FT n0 = FT(mp.rain.n0) * FT(1e-12);
    FT lam_inv = lambda_inverse<FT>(FT(mp.rain.n0), mp.rain.mass, q, rho) / FT(1e-3);
    FT Z = 720 * n0 * pow_(lam_inv, FT(7));
    FT log_10_Z0 = FT(-18);
    FT log_Z = FT(10) * (log10_(Z) - log_10_Z0 - FT(9));
    return jmax(FT(-150), log_Z);
}

template <class FT> inline bool diag_notvalid(FT B) { return B == FT(0) || !isfinite_(B); }

// CMD.radar_reflectivity_2M((; pdf_c, pdf_r)::SB2006, q_lcl, q_rai, N_lcl, N_rai, ρ_air)   CloudDiagnostics.jl:60-79
template <class FT>
inline FT radar_reflectivity_2M(const typename PT<FT>::sb_pdf_c& pdf_c, const typename PT<FT>::sb_pdf_r& pdf_r, FT q_lcl, FT q_rai,
                                FT N_lcl, FT N_rai, FT rho) {
    FT C = FT(4.0 / 3 * 3.141592653589793 * double(pdf_r.rho_w));
    FT Ar, Br, Ac, Bc;
    pdf_rain_parameters_mass<FT>(pdf_r, q_rai, rho, N_rai, Ar, Br);
    pdf_cloud_parameters_mass<FT>(pdf_c, q_lcl, rho, N_lcl, Ac, Bc);
    FT Zc = diag_notvalid(Bc) ? FT(0) : FT(generalized_gamma_Mn<FT>(pdf_c.nu_c, pdf_c.mu_c, Bc, N_lcl, FT(2)) / (C * C));
    FT Zr = diag_notvalid(Br) ? FT(0) : FT(generalized_gamma_Mn<FT>(pdf_r.nu_r, pdf_r.mu_r, Br, N_rai, FT(2)) / (C * C));
    return jmax(FT(-150), FT(10 * (log10_(jmax(FT(0), FT(Zc + Zr))) - FT(-18))));
}

// CMD.effective_radius_2M                                                           CloudDiagnostics.jl:95-116
template <class FT>
inline FT effective_radius_2M(const typename PT<FT>::sb_pdf_c& pdf_c, const typename PT<FT>::sb_pdf_r& pdf_r, FT q_lcl, FT q_rai,
                              FT N_lcl, FT N_rai, FT rho) {
    FT C = FT(4.0 / 3 * 3.141592653589793 * double(pdf_r.rho_w));
    FT Ar, Br, Ac, Bc;
    pdf_rain_parameters_mass<FT>(pdf_r, q_rai, rho, N_rai, Ar, Br);
    pdf_cloud_parameters_mass<FT>(pdf_c, q_lcl, rho, N_lcl, Ac, Bc);
    const bool nc = diag_notvalid(Bc), nr = diag_notvalid(Br);
    FT M3_c = nc ? FT(0) : FT(generalized_gamma_Mn<FT>(pdf_c.nu_c, pdf_c.mu_c, Bc, N_lcl, FT(1)) / C);
    FT M3_r = nr ? FT(0) : FT(generalized_gamma_Mn<FT>(pdf_r.nu_r, pdf_r.mu_r, Br, N_rai, FT(1)) / C);
    FT n_mass = FT(2) / 3;
    FT Cn = pow_(C, n_mass);
    FT M2_c = nc ? FT(0) : FT(generalized_gamma_Mn<FT>(pdf_c.nu_c, pdf_c.mu_c, Bc, N_lcl, n_mass) / Cn);
    FT M2_r = nr ? FT(0) : FT(generalized_gamma_Mn<FT>(pdf_r.nu_r, pdf_r.mu_r, Br, N_rai, n_mass) / Cn);
    return (M2_c + M2_r <= eps_numerics<FT>()) ? FT(0) : FT((M3_c + M3_r) / (M2_c + M2_r));
}

// CMD.effective_radius_Liu_Hallet_97((; ρw), ρ_air, q_lcl, N_lcl, q_rai, N_rai)      CloudDiagnostics.jl:132-148
template <class FT> inline FT effective_radius_Liu_Hallet_97(FT rho_w, FT rho, FT q_lcl, FT N_lcl, FT q_rai, FT N_rai) {
    FT k = FT(0.8);
    FT r_vol = ((N_lcl + N_rai) < eps_numerics<FT>())
                   ? FT(0)
                   : FT(pow_(FT((FT(3) * (q_lcl + q_rai) * rho) / (FT(4) * pi<FT>() * rho_w * (N_lcl + N_rai))), FT(1.0 / 3)));
    return r_vol / pow_(k, FT(1.0 / 3));
}

}  // namespace orc
