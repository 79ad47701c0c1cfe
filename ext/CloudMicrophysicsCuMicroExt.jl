"""
    CloudMicrophysicsCuMicroExt

Package extension of CloudMicrophysics.jl (weak dependency: CUDA.jl) that adds ARRAY methods
to the existing scalar API — same function names, same argument order — dispatching on
`CuVector`s.  Each method flattens the live parameter objects (`mp`, `tps`, `sb`, ...) field by
field into the plain-old-data blocks of `include/cumicro_params.inc` and `ccall`s one entry point
of `libcumicro.so` (`include/cumicro.h`), i.e. one fused sm_100a kernel per call.

The reference's own extension mechanism is used (`Project.toml:20-25`,
`ext/EmulatorModelsExt.jl:32-44`): add to the reference's Project.toml

    [weakdeps]
    CUDA = "052768ef-5323-5732-b1bb-66c8b64840ba"
    [extensions]
    CloudMicrophysicsCuMicroExt = "CUDA"

and copy this file to `ext/`.  `ENV["LIBCUMICRO"]` names the shared library (default
`libcumicro.so` on the loader path).

Layout contract.  Every `C*` struct below mirrors one C struct of `include/cumicro_params.inc`:
same field order, `FT` fields first, `Int32` fields last (so natural alignment equals the C
layout for both Float64 and Float32).  `tests/test_julia_ext.py` parses this file and the header
and fails if a mirror drifts (names, order, element types, array lengths); `tests/test_abi.py`
pins `sizeof`/`offsetof` of the C side against gcc.  A plain `reinterpret` of the reference's own
structs is NOT layout-safe (nested parametric types, `Nothing` slots, `Int` fields, NamedTuples).

Scalar methods are untouched: the array methods are strictly more specific (`CuVector` arguments).
"""
module CloudMicrophysicsCuMicroExt

import CUDA
import CUDA: CuVector, CuPtr, CU_NULL

import CloudMicrophysics.BulkMicrophysicsTendencies as BMT
import CloudMicrophysics.Microphysics1M as CM1
import CloudMicrophysics.Microphysics2M as CM2
import CloudMicrophysics.MicrophysicsNonEq as CMNonEq
import CloudMicrophysics.HetIceNucleation as CM_HetIce
import CloudMicrophysics.HomIceNucleation as CM_HomIce
import CloudMicrophysics.AerosolActivation as AA
import CloudMicrophysics.AerosolModel as AM
import CloudMicrophysics.CloudDiagnostics as CMD
import CloudMicrophysics.Common as CO
import CloudMicrophysics.P3Scheme as P3
import CloudMicrophysics.Quadrature as QUAD
import CloudMicrophysics.Parameters as CMP
import CloudMicrophysics.ThermodynamicsInterface as TDI

const TDP = TDI.TD.Parameters
const libcumicro = get(ENV, "LIBCUMICRO", "libcumicro.so")

# =========================================================================================
# POD mirrors of include/cumicro_params.inc   (BEGIN-MIRRORS: parsed by tests/test_julia_ext.py)
# =========================================================================================

# cumicro_thermo
struct CThermo{FT}
    T_0::FT
    T_triple::FT
    press_triple::FT
    T_freeze::FT
    R_v::FT
    R_d::FT
    cp_d::FT
    cp_v::FT
    cp_l::FT
    cp_i::FT
    LH_v0::FT
    LH_s0::FT
    q_min::FT
    grav::FT
end

# cumicro_air
struct CAir{FT}
    K_therm::FT
    D_vapor::FT
    nu_air::FT
end

# cumicro_sb_pdf_c
struct CSbPdfC{FT}
    nu_c::FT
    mu_c::FT
    xc_min::FT
    xc_max::FT
    rho_w::FT
    loggamma_z1::FT
    loggamma_z2::FT
end

# cumicro_sb_pdf_r
struct CSbPdfR{FT}
    nu_r::FT
    mu_r::FT
    xr_min::FT
    xr_max::FT
    N0_min::FT
    N0_max::FT
    lam_min::FT
    lam_max::FT
    rho_w::FT
    rho0::FT
    limited::Int32
    _pad::Int32
end

# cumicro_sb_acnv
struct CSbAcnv{FT}
    kcc::FT
    x_star::FT
    rho0::FT
    A::FT
    a::FT
    b::FT
end

# cumicro_sb_accr
struct CSbAccr{FT}
    kcr::FT
    tau0::FT
    rho0::FT
    c::FT
end

# cumicro_sb_self
struct CSbSelf{FT}
    krr::FT
    kappa_rr::FT
    d::FT
end

# cumicro_sb_brek
struct CSbBrek{FT}
    Deq::FT
    Dr_th::FT
    kbr::FT
    kappa_br::FT
end

# cumicro_sb_evap
struct CSbEvap{FT}
    av::FT
    bv::FT
    alpha::FT
    beta::FT
    rho0::FT
    a_vent_1::FT
    b_vent_1::FT
    a_vent_0_coeff::FT
    b_vent_0_coeff::FT
    beta_vent_0::FT
end

# cumicro_sb2006
struct CSb2006{FT}
    pdf_c::CSbPdfC{FT}
    pdf_r::CSbPdfR{FT}
    acnv::CSbAcnv{FT}
    accr::CSbAccr{FT}
    self::CSbSelf{FT}
    brek::CSbBrek{FT}
    evap::CSbEvap{FT}
    numadj_tau::FT
end

# cumicro_vel_sb2006
struct CVelSb2006{FT}
    rho0::FT
    aR::FT
    bR::FT
    cR::FT
end

# cumicro_vel_stokes
struct CVelStokes{FT}
    rho_w::FT
    nu_air::FT
    grav::FT
end

# cumicro_vel_chen_rain
struct CVelChenRain{FT}
    rho0::FT
    a::NTuple{3, FT}
    a3_pow::FT
    b::NTuple{3, FT}
    b_rho::FT
    c::NTuple{3, FT}
end

# cumicro_vel_chen_small_ice
struct CVelChenSmallIce{FT}
    A::NTuple{3, FT}
    B::NTuple{3, FT}
    C::NTuple{4, FT}
    E::NTuple{3, FT}
    F::NTuple{3, FT}
    G::NTuple{3, FT}
    cutoff::FT
end

# cumicro_vel_chen_large_ice
struct CVelChenLargeIce{FT}
    A::NTuple{3, FT}
    B::NTuple{3, FT}
    C::NTuple{3, FT}
    E::NTuple{3, FT}
    F::NTuple{3, FT}
    G::NTuple{3, FT}
    H::NTuple{3, FT}
    cutoff::FT
end

# cumicro_params_2m_warm
struct CParams2mWarm{FT}
    tps::CThermo{FT}
    sb::CSb2006{FT}
    aps::CAir{FT}
    condevap_tau_relax::FT
    subdep_tau_relax::FT
end

# cumicro_particle_mass
struct CParticleMass{FT}
    r0::FT
    m0::FT
    me::FT
    dm::FT
    chi_m::FT
    gamma_coeff::FT
end

# cumicro_particle_area
struct CParticleArea{FT}
    a0::FT
    ae::FT
    da::FT
    chi_a::FT
end

# cumicro_ventilation
struct CVentilation{FT}
    a::FT
    b::FT
end

# cumicro_cloud_liquid
struct CCloudLiquid{FT}
    rho_w::FT
    r_eff::FT
    N_0::FT
end

# cumicro_cloud_ice
struct CCloudIce{FT}
    n0::FT
    mass::CParticleMass{FT}
    rho_i::FT
    r_eff::FT
    N_0::FT
end

# cumicro_rain
struct CRain{FT}
    n0::FT
    mass::CParticleMass{FT}
    area::CParticleArea{FT}
    vent::CVentilation{FT}
end

# cumicro_snow
struct CSnow{FT}
    mu::FT
    nu::FT
    mass::CParticleMass{FT}
    area::CParticleArea{FT}
    vent::CVentilation{FT}
    aspr_phi::FT
    aspr_kappa::FT
    rho_i::FT
    gamma_aspect_oblate::FT
    gamma_aspect_prolate::FT
end

# cumicro_vel_blk1m_rain
struct CVelBlk1mRain{FT}
    r0::FT
    ve::FT
    dv::FT
    chi_v::FT
    rho_w::FT
    C_drag::FT
    grav::FT
    gamma_vent::FT
    gamma_term::FT
    gamma_accr::FT
    gamma_accr_rain_sink::FT
end

# cumicro_vel_blk1m_snow
struct CVelBlk1mSnow{FT}
    r0::FT
    ve::FT
    dv::FT
    chi_v::FT
    v0::FT
    gamma_vent::FT
    gamma_term::FT
    gamma_accr::FT
end

# cumicro_frostenberg2023
struct CFrostenberg2023{FT}
    sigma::FT
    a::FT
    b::FT
    T_freeze::FT
    log_a::FT
end

# cumicro_options_1m
struct COptions1m
    cloud_liquid_formation::Int32
    cloud_ice_formation::Int32
    cloud_ice_melt::Int32
    rain_autoconversion::Int32
    snow_autoconversion::Int32
    rain_condensation_evaporation::Int32
    snow_deposition_sublimation::Int32
    snow_melt::Int32
    cloud_liquid_rain_accretion::Int32
    cloud_liquid_snow_accretion::Int32
    cloud_ice_rain_accretion::Int32
    cloud_ice_snow_accretion::Int32
    rain_snow_accretion::Int32
    _pad::Int32
end

# cumicro_process_params_1m
struct CProcessParams1m{FT}
    cloud_liquid_tau_relax::FT
    cloud_ice_tau_relax::FT
    frostenberg::CFrostenberg2023{FT}
    rain_acnv_tau::FT
    rain_acnv_q_threshold::FT
    rain_acnv_k::FT
    rain_acnv_alpha::FT
    rain_acnv_Nc::FT
    snow_acnv_tau::FT
    snow_acnv_q_threshold::FT
    snow_acnv_k::FT
    snow_acnv_r_ice_snow::FT
    e_lcl_rai::FT
    e_lcl_sno::FT
    e_icl_rai::FT
    e_icl_sno::FT
    e_rai_sno::FT
    coeff_disp::FT
end

# cumicro_params_1m
struct CParams1m{FT}
    tps::CThermo{FT}
    cloud_liquid::CCloudLiquid{FT}
    cloud_ice::CCloudIce{FT}
    rain::CRain{FT}
    snow::CSnow{FT}
    aps::CAir{FT}
    vel_rain::CVelBlk1mRain{FT}
    vel_snow::CVelBlk1mSnow{FT}
    pp::CProcessParams1m{FT}
    processes::COptions1m
end

# cumicro_dust
struct CDust{FT}
    deposition_m::FT
    deposition_c::FT
    ABIFM_m::FT
    ABIFM_c::FT
    S0_warm::FT
    S0_cold::FT
    a_warm::FT
    a_cold::FT
    has_deposition::Int32
    has_ABIFM::Int32
end

# cumicro_koop2000
struct CKoop2000{FT}
    da_w_min::FT
    da_w_max::FT
    c1::FT
    c2::FT
    c3::FT
    c4::FT
    linear_c1::FT
    linear_c2::FT
end

# cumicro_mohler2006
struct CMohler2006{FT}
    Si_max::FT
    T_thr::FT
end

# cumicro_mm2014
struct CMm2014{FT}
    c1::FT
    c2::FT
    T0::FT
    T_dep_thres::FT
    het_a::FT
    het_B::FT
end

# cumicro_h2so4
struct CH2so4{FT}
    T_max::FT
    T_min::FT
    w_2::FT
    c::NTuple{7, FT}
end

# cumicro_arg2000
struct CArg2000{FT}
    M_w::FT
    R::FT
    rho_w::FT
    rho_i::FT
    sigma::FT
    g::FT
    f1::FT
    f2::FT
    g1::FT
    g2::FT
    p1::FT
    p2::FT
end

# cumicro_aerosol_mode
struct CAerosolMode{FT}
    r_dry::FT
    stdev::FT
    N::FT
    hygro::FT
    molar_mass_mix::FT
end

# cumicro_params_icenuc
struct CParamsIcenuc{FT}
    tps::CThermo{FT}
    aps::CAir{FT}
    arg::CArg2000{FT}
    dust::CDust{FT}
    koop::CKoop2000{FT}
    mohler::CMohler2006{FT}
    mm2014::CMm2014{FT}
    h2so4::CH2so4{FT}
    frostenberg::CFrostenberg2023{FT}
    modes::NTuple{8, CAerosolMode{FT}}
    n_modes::Int32
    hom_linear::Int32
end

# cumicro_p3_scheme
struct CP3Scheme{FT}
    alpha_va::FT
    beta_va::FT
    gamma::FT
    sigma::FT
    slope_a::FT
    slope_b::FT
    slope_c::FT
    slope_mu_max::FT
    slope_mu_const::FT
    vent_a::FT
    vent_b::FT
    rim_a::FT
    rim_b::FT
    rim_c::FT
    rim_rho_ice::FT
    tau_wet::FT
    rho_i::FT
    rho_l::FT
    T_freeze::FT
    slope_power_law::Int32
    aspect_oblate::Int32
end

# cumicro_quadrature
struct CQuadrature{FT}
    nodes::NTuple{128, FT}
    weights::NTuple{128, FT}
    n::Int32
    gauss_legendre::Int32
end

# cumicro_params_p3
struct CParamsP3{FT}
    warm::CParams2mWarm{FT}
    scheme::CP3Scheme{FT}
    vel_rain::CVelChenRain{FT}
    vel_small_ice::CVelChenSmallIce{FT}
    vel_large_ice::CVelChenLargeIce{FT}
    ice_nucleation::CFrostenberg2023{FT}
    rain_freezing_het_a::FT
    rain_freezing_het_B::FT
    tau_act::FT
    quad::CQuadrature{FT}
end

# cumicro_params_0m
struct CParams0m{FT}
    tau_precip::FT
    qc_0::FT
    S_0::FT
end

# cumicro_params_2m_alt
struct CParams2mAlt{FT}
    kk_acnv_A::FT
    kk_acnv_a::FT
    kk_acnv_b::FT
    kk_acnv_c::FT
    kk_accr_A::FT
    kk_accr_a::FT
    kk_accr_b::FT
    b_acnv_C::FT
    b_acnv_a::FT
    b_acnv_b::FT
    b_acnv_c::FT
    b_acnv_N_0::FT
    b_acnv_k::FT
    b_acnv_d_low::FT
    b_acnv_d_high::FT
    b_accr_A::FT
    tc_acnv_m0_liq_coeff::FT
    tc_acnv_me_liq::FT
    tc_acnv_D::FT
    tc_acnv_a::FT
    tc_acnv_b::FT
    tc_acnv_r_0::FT
    tc_acnv_k::FT
    tc_accr_A::FT
    ld_rho_w::FT
    ld_R_6C_0::FT
    ld_E_0::FT
    ld_k::FT
end

# cumicro_params_emulator
struct CParamsEmulator{FT}
    mode_N::NTuple{8, FT}
    mode_mean::NTuple{8, FT}
    mode_stdev::NTuple{8, FT}
    mode_kappa::NTuple{8, FT}
    feat_mean::NTuple{35, FT}
    feat_inv_scale::NTuple{35, FT}
    n_modes::Int32
    n_layers::Int32
    width::NTuple{4, Int32}
    activation::Int32
    log_features::Int32
    target_transform::Int32
end

# (END-MIRRORS)

# =========================================================================================
# Packers: live parameter objects -> POD blocks, field by field
# =========================================================================================

pack_thermo(::Type{FT}, tps) where {FT} = CThermo{FT}(
    TDP.T_0(tps), TDP.T_triple(tps), TDP.press_triple(tps), TDP.T_freeze(tps), TDP.R_v(tps), TDP.R_d(tps),
    TDP.cp_d(tps), TDP.cp_v(tps), TDP.cp_l(tps), TDP.cp_i(tps), TDP.LH_v0(tps), TDP.LH_s0(tps),
    TDP.q_min(tps), TDP.grav(tps),
)

pack_air(::Type{FT}, a::CMP.AirProperties) where {FT} = CAir{FT}(a.K_therm, a.D_vapor, a.ν_air)

pack_pdf_c(::Type{FT}, p::CMP.CloudParticlePDF_SB2006) where {FT} =
    CSbPdfC{FT}(p.νc, p.μc, p.xc_min, p.xc_max, p.ρw, p.loggamma_z1, p.loggamma_z2)

pack_pdf_r(::Type{FT}, p::CMP.RainParticlePDF_SB2006_limited) where {FT} =
    CSbPdfR{FT}(p.νr, p.μr, p.xr_min, p.xr_max, p.N0_min, p.N0_max, p.λ_min, p.λ_max, p.ρw, p.ρ0, Int32(1), Int32(0))
# the not-limited variant has no limiter fields: zeros, limited = 0 (they are never read)
pack_pdf_r(::Type{FT}, p::CMP.RainParticlePDF_SB2006_notlimited) where {FT} =
    CSbPdfR{FT}(p.νr, p.μr, p.xr_min, p.xr_max, 0, 0, 0, 0, p.ρw, p.ρ0, Int32(0), Int32(0))

pack_acnv(::Type{FT}, p::CMP.AcnvSB2006) where {FT} = CSbAcnv{FT}(p.kcc, p.x_star, p.ρ0, p.A, p.a, p.b)
pack_accr(::Type{FT}, p::CMP.AccrSB2006) where {FT} = CSbAccr{FT}(p.kcr, p.τ0, p.ρ0, p.c)
pack_self(::Type{FT}, p::CMP.SelfColSB2006) where {FT} = CSbSelf{FT}(p.krr, p.κrr, p.d)
pack_brek(::Type{FT}, p::CMP.BreakupSB2006) where {FT} = CSbBrek{FT}(p.Deq, p.Dr_th, p.kbr, p.κbr)
pack_evap(::Type{FT}, p::CMP.EvaporationSB2006) where {FT} = CSbEvap{FT}(
    p.av, p.bv, p.α, p.β, p.ρ0, p.a_vent_1, p.b_vent_1, p.a_vent_0_coeff, p.b_vent_0_coeff, p.β_vent_0,
)

pack_sb2006(::Type{FT}, sb::CMP.SB2006) where {FT} = CSb2006{FT}(
    pack_pdf_c(FT, sb.pdf_c), pack_pdf_r(FT, sb.pdf_r), pack_acnv(FT, sb.acnv), pack_accr(FT, sb.accr),
    pack_self(FT, sb.self), pack_brek(FT, sb.brek), pack_evap(FT, sb.evap), sb.numadj.τ,
)

pack_vel(::Type{FT}, v::CMP.SB2006VelType) where {FT} = CVelSb2006{FT}(v.ρ0, v.aR, v.bR, v.cR)
pack_vel(::Type{FT}, v::CMP.StokesRegimeVelType) where {FT} = CVelStokes{FT}(v.ρw, v.ν_air, v.grav)
pack_vel(::Type{FT}, v::CMP.Chen2022VelTypeRain) where {FT} =
    CVelChenRain{FT}(v.ρ0, FT.(v.a), v.a3_pow, FT.(v.b), v.b_ρ, FT.(v.c))
pack_vel(::Type{FT}, v::CMP.Chen2022VelTypeSmallIce) where {FT} =
    CVelChenSmallIce{FT}(FT.(v.A), FT.(v.B), FT.(v.C), FT.(v.E), FT.(v.F), FT.(v.G), v.cutoff)
pack_vel(::Type{FT}, v::CMP.Chen2022VelTypeLargeIce) where {FT} =
    CVelChenLargeIce{FT}(FT.(v.A), FT.(v.B), FT.(v.C), FT.(v.E), FT.(v.F), FT.(v.G), FT.(v.H), v.cutoff)
pack_vel(::Type{FT}, v::CMP.Blk1MVelTypeRain) where {FT} = CVelBlk1mRain{FT}(
    v.r0, v.ve, v.Δv, v.χv, v.ρw, v.C_drag, v.grav, v.gamma_vent, v.gamma_term, v.gamma_accr, v.gamma_accr_rain_sink,
)
pack_vel(::Type{FT}, v::CMP.Blk1MVelTypeSnow) where {FT} =
    CVelBlk1mSnow{FT}(v.r0, v.ve, v.Δv, v.χv, v.v0, v.gamma_vent, v.gamma_term, v.gamma_accr)

function pack_warm(::Type{FT}, wr::CMP.WarmRainParams2M, tps) where {FT}
    return CParams2mWarm{FT}(
        pack_thermo(FT, tps), pack_sb2006(FT, wr.seifert_beheng), pack_air(FT, wr.air_properties),
        wr.condevap.τ_relax, wr.subdep.τ_relax,
    )
end
pack(::Type{FT}, mp::CMP.Microphysics2MParams{WR, Nothing}, tps) where {FT, WR} = pack_warm(FT, mp.warm_rain, tps)

# ---- 1-moment ---------------------------------------------------------------------------
pack_mass(::Type{FT}, m::CMP.ParticleMass) where {FT} = CParticleMass{FT}(m.r0, m.m0, m.me, m.Δm, m.χm, m.gamma_coeff)
pack_area(::Type{FT}, a::CMP.ParticleArea) where {FT} = CParticleArea{FT}(a.a0, a.ae, a.Δa, a.χa)
pack_vent(::Type{FT}, v::CMP.Ventilation) where {FT} = CVentilation{FT}(v.a, v.b)
pack_frostenberg(::Type{FT}, f::CMP.Frostenberg2023) where {FT} = CFrostenberg2023{FT}(f.σ, f.a, f.b, f.T_freeze, f.log_a)
pack_frostenberg(::Type{FT}, ::Nothing) where {FT} = CFrostenberg2023{FT}(0, 0, 0, 0, 0)

# Microphysics1MOptions: `nothing` -> 0, the variants in declaration order -> 1, 2   (CUMICRO_1M_* of cumicro.h)
opt_code(::Nothing) = Int32(0)
opt_code(::CMP.MicrophysicsOption) = Int32(1)
opt_code(::CMP.TemperatureDependent) = Int32(2)
opt_code(::CMP.PrescribedNd) = Int32(2)
opt_code(::CMP.WithSupersaturation) = Int32(2)
opt_code(::CMP.DepositionAndSublimation) = Int32(2)

pack_options(o::CMP.Microphysics1MOptions) = COptions1m(
    opt_code(o.cloud_liquid_formation), opt_code(o.cloud_ice_formation), opt_code(o.cloud_ice_melt),
    opt_code(o.rain_autoconversion), opt_code(o.snow_autoconversion), opt_code(o.rain_condensation_evaporation),
    opt_code(o.snow_deposition_sublimation), opt_code(o.snow_melt), opt_code(o.cloud_liquid_rain_accretion),
    opt_code(o.cloud_liquid_snow_accretion), opt_code(o.cloud_ice_rain_accretion), opt_code(o.cloud_ice_snow_accretion),
    opt_code(o.rain_snow_accretion), Int32(0),
)

# process_params (Microphysics1MOptions.jl:301-396): a NamedTuple with one entry per process; entries of disabled
# processes are `nothing`, entries of other variants do not exist.  Missing members pack as zero (never read).
_get(x, name::Symbol, ::Type{FT}) where {FT} = (x !== nothing && hasproperty(x, name)) ? FT(getproperty(x, name)) : zero(FT)
function pack_process_params(::Type{FT}, o::CMP.Microphysics1MOptions, pp) where {FT}
    cif = pp.cloud_ice_formation
    fr = (cif !== nothing && hasproperty(cif, :frostenberg)) ? cif.frostenberg : nothing
    ra, sa = pp.rain_autoconversion, pp.snow_autoconversion
    return CProcessParams1m{FT}(
        _get(pp.cloud_liquid_formation, :τ_relax, FT),
        _get(cif, :τ_relax, FT),
        pack_frostenberg(FT, fr),
        _get(ra, :τ, FT), _get(ra, :q_threshold, FT), _get(ra, :k, FT),       # Kessler1M: Acnv1M{τ, q_threshold, k}
        _get(ra, :α, FT), _get(ra, :Nc, FT),                                   # PrescribedNd: VarTimescaleAcnv{τ, α, Nc}
        _get(sa, :τ, FT), _get(sa, :q_threshold, FT), _get(sa, :k, FT),       # NoSupersaturation: Acnv1M
        _get(sa, :r_ice_snow, FT),                                             # WithSupersaturation
        _get(pp.cloud_liquid_rain_accretion, :e, FT), _get(pp.cloud_liquid_snow_accretion, :e, FT),
        _get(pp.cloud_ice_rain_accretion, :e, FT), _get(pp.cloud_ice_snow_accretion, :e, FT),
        _get(pp.rain_snow_accretion, :e, FT), _get(pp.rain_snow_accretion, :coeff_disp, FT),
    )
end

function pack(::Type{FT}, mp::CMP.Microphysics1MParams, tps) where {FT}
    liq, ice = mp.cloud.liquid, mp.cloud.ice
    rain, snow = mp.precip.rain, mp.precip.snow
    vel = mp.terminal_velocity          # Blk1MVelType{rain, snow}
    return CParams1m{FT}(
        pack_thermo(FT, tps),
        CCloudLiquid{FT}(liq.ρw, liq.r_eff, liq.N_0),
        CCloudIce{FT}(ice.pdf.n0, pack_mass(FT, ice.mass), ice.ρᵢ, ice.r_eff, ice.N_0),
        CRain{FT}(rain.pdf.n0, pack_mass(FT, rain.mass), pack_area(FT, rain.area), pack_vent(FT, rain.vent)),
        CSnow{FT}(
            snow.pdf.μ, snow.pdf.ν, pack_mass(FT, snow.mass), pack_area(FT, snow.area), pack_vent(FT, snow.vent),
            snow.aspr.ϕ, snow.aspr.κ, snow.ρᵢ, snow.gamma_aspect_oblate, snow.gamma_aspect_prolate,
        ),
        pack_air(FT, mp.air_properties),
        pack_vel(FT, vel.rain), pack_vel(FT, vel.snow),
        pack_process_params(FT, mp.processes, mp.process_params),
        pack_options(mp.processes),
    )
end

pack(::Type{FT}, mp::CMP.Microphysics0MParams, tps) where {FT} = CParams0m{FT}(mp.precip.τ_precip, mp.precip.qc_0, mp.precip.S_0)

# ---- ice nucleation / aerosol activation ------------------------------------------------------
# The reference returns zero for aerosol types without a parameterisation (IN:102, IN:134): has_* = 0.
function pack_dust(::Type{FT}, d::CMP.AerosolType) where {FT}
    g(name) = hasproperty(d, name) ? FT(getproperty(d, name)) : zero(FT)
    return CDust{FT}(
        g(:deposition_m), g(:deposition_c), g(:ABIFM_m), g(:ABIFM_c), g(:S₀_warm), g(:S₀_cold), g(:a_warm), g(:a_cold),
        Int32(hasproperty(d, :deposition_m)), Int32(hasproperty(d, :ABIFM_m)),
    )
end
pack_dust(::Type{FT}, ::Nothing) where {FT} = CDust{FT}(0, 0, 0, 0, 0, 0, 0, 0, Int32(0), Int32(0))
pack_koop(::Type{FT}, k::CMP.Koop2000) where {FT} =
    CKoop2000{FT}(k.Δa_w_min, k.Δa_w_max, k.c₁, k.c₂, k.c₃, k.c₄, k.linear_c₁, k.linear_c₂)
pack_koop(::Type{FT}, ::Nothing) where {FT} = CKoop2000{FT}(0, 0, 0, 0, 0, 0, 0, 0)
pack_mohler(::Type{FT}, m::CMP.Mohler2006) where {FT} = CMohler2006{FT}(m.Sᵢ_max, m.T_thr)
pack_mohler(::Type{FT}, ::Nothing) where {FT} = CMohler2006{FT}(0, 0)
pack_mm2014(::Type{FT}, m::CMP.MorrisonMilbrandt2014) where {FT} = CMm2014{FT}(m.c₁, m.c₂, m.T₀, m.T_dep_thres, m.het_a, m.het_B)
pack_mm2014(::Type{FT}, ::Nothing) where {FT} = CMm2014{FT}(0, 0, 0, 0, 0, 0)
pack_h2so4(::Type{FT}, h::CMP.H2SO4SolutionParameters) where {FT} =
    CH2so4{FT}(h.T_max, h.T_min, h.w_2, FT.((h.c1, h.c2, h.c3, h.c4, h.c5, h.c6, h.c7)))
pack_h2so4(::Type{FT}, ::Nothing) where {FT} = CH2so4{FT}(0, 0, 0, ntuple(_ -> zero(FT), 7))
pack_arg(::Type{FT}, a::CMP.AerosolActivationParameters) where {FT} =
    CArg2000{FT}(a.M_w, a.R, a.ρ_w, a.ρ_i, a.σ, a.g, a.f1, a.f2, a.g1, a.g2, a.p1, a.p2)
pack_arg(::Type{FT}, ::Nothing) where {FT} = CArg2000{FT}(ntuple(_ -> zero(FT), 12)...)

# Modes: the per-mode mean hygroscopicity (AA:55-95) and Σ molar_mass mass_mix_ratio (AA:318) depend on parameters only.
function pack_modes(::Type{FT}, ap, ad) where {FT}
    n = AM.n_modes(ad)
    n <= 8 || throw(ArgumentError("cumicro: at most 8 aerosol modes (got $n)"))
    hyg = AA.mean_hygroscopicity_parameter(ap, ad)
    zero_mode = CAerosolMode{FT}(0, 0, 0, 0, 0)
    modes = ntuple(8) do i
        i > n && return zero_mode
        m = ad.modes[i]
        mm = sum(FT(m.molar_mass[j]) * FT(m.mass_mix_ratio[j]) for j in 1:AM.n_components(m))
        CAerosolMode{FT}(m.r_dry, m.stdev, m.N, hyg[i], mm)
    end
    return modes, Int32(n)
end

function pack_icenuc(::Type{FT}, tps; aps = nothing, ap = nothing, ad = nothing, dust = nothing, koop = nothing,
                     mohler = nothing, mm2014 = nothing, h2so4 = nothing, frostenberg = nothing, hom_linear = false) where {FT}
    modes, n_modes = (ap === nothing || ad === nothing) ? (ntuple(_ -> CAerosolMode{FT}(0, 0, 0, 0, 0), 8), Int32(0)) :
                     pack_modes(FT, ap, ad)
    return CParamsIcenuc{FT}(
        pack_thermo(FT, tps), aps === nothing ? CAir{FT}(0, 0, 0) : pack_air(FT, aps), pack_arg(FT, ap), pack_dust(FT, dust),
        pack_koop(FT, koop), pack_mohler(FT, mohler), pack_mm2014(FT, mm2014), pack_h2so4(FT, h2so4),
        pack_frostenberg(FT, frostenberg), modes, n_modes, Int32(hom_linear),
    )
end

# ---- P3 ---------------------------------------------------------------------------------------
function pack_scheme(::Type{FT}, p::CMP.ParametersP3) where {FT}
    sl = p.slope
    pl = sl isa CMP.SlopePowerLaw
    return CP3Scheme{FT}(
        p.mass.α_va, p.mass.β_va, p.area.γ, p.area.σ,
        pl ? sl.a : zero(FT), pl ? sl.b : zero(FT), pl ? sl.c : zero(FT), pl ? sl.μ_max : zero(FT), pl ? zero(FT) : sl.μ,
        p.vent.aᵥ, p.vent.bᵥ, p.ρ_rim_local.a, p.ρ_rim_local.b, p.ρ_rim_local.c, p.ρ_rim_local.ρ_ice,
        p.τ_wet, p.ρ_i, p.ρ_l, p.T_freeze, Int32(pl), Int32(p.aspect_ratio isa CMP.Oblate),
    )
end

# Quadrature.jl:166-236.  ChebyshevGauss: node y_i = cospi((2i-1)/(2n)), total weight sqrt(1-y_i²) π/n, in FT as `integrate` forms them.
function pack_quad(::Type{FT}, q::QUAD.GaussLegendre) where {FT}
    n = q.n
    n <= 128 || throw(ArgumentError("cumicro: quadrature order $n > 128"))
    nodes = ntuple(i -> i <= n ? FT(q.nodes[i]) : zero(FT), 128)
    weights = ntuple(i -> i <= n ? FT(q.weights[i]) : zero(FT), 128)
    return CQuadrature{FT}(nodes, weights, Int32(n), Int32(1))
end
function pack_quad(::Type{FT}, q::QUAD.ChebyshevGauss) where {FT}
    n = q.n
    n <= 128 || throw(ArgumentError("cumicro: quadrature order $n > 128"))
    y(i) = cospi((2 * FT(i) - 1) / (2n))
    nodes = ntuple(i -> i <= n ? FT(y(i)) : zero(FT), 128)
    weights = ntuple(i -> i <= n ? FT(sqrt(1 - y(i)^2) * (FT(π) / n)) : zero(FT), 128)
    return CQuadrature{FT}(nodes, weights, Int32(n), Int32(0))
end

function pack(::Type{FT}, mp::CMP.Microphysics2MParams{WR, ICE}, tps; quad = mp.ice.quad) where {FT, WR, ICE <: CMP.P3IceParams}
    ice = mp.ice
    sb = mp.warm_rain.seifert_beheng
    # BMT:953-955 reads the P3-side copies of the cloud / rain PSD; the block carries one copy
    (ice.cloud_pdf == sb.pdf_c && ice.rain_pdf == sb.pdf_r) ||
        throw(ArgumentError("cumicro: mp.ice.cloud_pdf / rain_pdf differ from mp.warm_rain.seifert_beheng.pdf_c / pdf_r"))
    vel = ice.terminal_velocity          # Chen2022VelType{rain, small_ice, large_ice}
    return CParamsP3{FT}(
        pack_warm(FT, mp.warm_rain, tps), pack_scheme(FT, ice.scheme), pack_vel(FT, vel.rain), pack_vel(FT, vel.small_ice),
        pack_vel(FT, vel.large_ice), pack_frostenberg(FT, ice.ice_nucleation), ice.rain_freezing.het_a, ice.rain_freezing.het_B,
        ice.inp_depletion_model.τ_act, pack_quad(FT, quad),
    )
end

# ---- alternative 2-moment closures ------------------------------------------------------------
function pack_alt(::Type{FT}, s) where {FT}
    z = zero(FT)
    kk = s isa CMP.KK2000 ? (s.acnv.A, s.acnv.a, s.acnv.b, s.acnv.c, s.accr.A, s.accr.a, s.accr.b) : ntuple(_ -> z, 7)
    b = s isa CMP.B1994 ? (s.acnv.C, s.acnv.a, s.acnv.b, s.acnv.c, s.acnv.N_0, s.acnv.k, s.acnv.d_low, s.acnv.d_high, s.accr.A) :
        ntuple(_ -> z, 9)
    tc = s isa CMP.TC1980 ? (s.acnv.m0_liq_coeff, s.acnv.me_liq, s.acnv.D, s.acnv.a, s.acnv.b, s.acnv.r_0, s.acnv.k, s.accr.A) :
         ntuple(_ -> z, 8)
    ld = s isa CMP.LD2004 ? (s.ρ_w, s.R_6C_0, s.E_0, s.k) : ntuple(_ -> z, 4)
    return CParams2mAlt{FT}(FT.(kk)..., FT.(b)..., FT.(tc)..., FT.(ld)...)
end

# =========================================================================================
# Call plumbing
# =========================================================================================

struct CuMicroError <: Exception
    status::Int
    msg::String
end
Base.showerror(io::IO, e::CuMicroError) = print(io, "cumicro: status ", e.status, ": ", e.msg)

last_error() = unsafe_string(ccall((:cumicro_last_error, libcumicro), Cstring, ()))
# <0: API misuse, caught before any launch -> ArgumentError; >0: a cudaError_t.
function check(st::Integer)
    st == 0 && return nothing
    st < 0 && throw(ArgumentError("cumicro ($st): " * last_error()))
    throw(CuMicroError(Int(st), last_error()))
end

sym(base::Symbol, ::Type{Float64}) = Symbol(base, :_f64)
sym(base::Symbol, ::Type{Float32}) = Symbol(base, :_f32)
cur_stream() = CUDA.stream().handle
const FTs = Union{Float32, Float64}
const Col{FT} = CuVector{FT}
# HOST table of device column pointers (the `FT* const*` arguments); `nothing` -> NULL (column skipped)
ptr_table(::Type{FT}, cols) where {FT} = CuPtr{FT}[c === nothing ? CuPtr{FT}(0) : pointer(c) for c in cols]
dev(::Type{FT}, c::CuVector{FT}) where {FT} = pointer(c)
dev(::Type{FT}, ::Nothing) where {FT} = CuPtr{FT}(0)
function same_length(cols...)
    n = length(first(cols))
    all(c -> c === nothing || length(c) == n, cols) || throw(DimensionMismatch("cumicro: columns differ in length"))
    return n
end

# Per-point domain violations cannot throw from a kernel: the kernels write NaN and bump a device counter.
# The wrapper re-raises the reference's exception.
function with_domain_counter(f, exc)
    counter = CUDA.zeros(UInt64, 1)
    out = f(pointer(counter))
    n = Array(counter)[1]          # synchronises the stream
    n == 0 || throw(exc(n))
    return out
end

# =========================================================================================
# BulkMicrophysicsTendencies
# =========================================================================================

# ---- 2-moment warm rain (BMT:820-854) -------------------------------------------------------
function BMT.bulk_microphysics_tendencies(
    ::BMT.Microphysics2Moment, mp::CMP.Microphysics2MParams{WR, Nothing}, tps,
    ρ::Col{FT}, T::Col{FT}, q_tot::Col{FT}, q_lcl::Col{FT}, n_lcl::Col{FT}, q_rai::Col{FT}, n_rai::Col{FT},
    q_ice::Union{Col{FT}, Nothing} = nothing,
) where {WR, FT <: FTs}
    n = same_length(ρ, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice)
    out = ntuple(_ -> similar(ρ), 4)
    zero4 = ntuple(_ -> similar(ρ), 4)            # the reference returns literal zeros for these four (BMT:840-853)
    blk = Ref(pack(FT, mp, tps))
    ztab = ptr_table(FT, zero4)
    GC.@preserve blk ztab begin
        st = ccall((sym(:cumicro_bmt2m_warm, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT},
             CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Ptr{CuPtr{FT}}, Ptr{Cvoid}),
            blk, n, ρ, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, dev(FT, q_ice),
            out[1], out[2], out[3], out[4], ztab, cur_stream())
    end
    check(st)
    return (; dq_lcl_dt = out[1], dn_lcl_dt = out[2], dq_rai_dt = out[3], dn_rai_dt = out[4],
            dq_ice_dt = zero4[1], dq_rim_dt = zero4[2], db_rim_dt = zero4[3], dn_lcl_activation_dt = zero4[4])
end

# ---- 2-moment + P3 (BMT:898-1083) --------------------------------------------------------------
function BMT.bulk_microphysics_tendencies(
    ::BMT.Microphysics2Moment, mp::CMP.Microphysics2MParams{WR, ICE}, tps,
    ρ::Col{FT}, T::Col{FT}, q_tot::Col{FT}, q_lcl::Col{FT}, n_lcl::Col{FT}, q_rai::Col{FT}, n_rai::Col{FT},
    q_ice::Col{FT}, n_ice::Col{FT}, q_rim::Col{FT}, b_rim::Col{FT}, logλ::Union{Col{FT}, Nothing},   # nothing: logλ is solved in the kernel (§8(f)-1)
    inpc_log_shift::Union{Col{FT}, Nothing} = nothing,
) where {WR, ICE <: CMP.P3IceParams, FT <: FTs}
    ins = (ρ, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, q_rim, b_rim, logλ)
    n = same_length(ins..., inpc_log_shift)
    out = ntuple(_ -> similar(ρ), 9)
    blk = Ref(pack(FT, mp, tps))
    itab, otab = ptr_table(FT, ins), ptr_table(FT, out)
    GC.@preserve blk itab otab begin
        st = ccall((sym(:cumicro_bmt2m_p3, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Int64, Ptr{CuPtr{FT}}, CuPtr{FT}, Ptr{CuPtr{FT}}, Ptr{Cvoid}),
            blk, n, itab, dev(FT, inpc_log_shift), otab, cur_stream())
    end
    check(st)
    return (; dq_lcl_dt = out[1], dn_lcl_dt = out[2], dq_rai_dt = out[3], dn_rai_dt = out[4], dq_ice_dt = out[5],
            dn_ice_dt = out[6], dq_rim_dt = out[7], db_rim_dt = out[8], dn_lcl_activation_dt = out[9])
end

# ---- 1-moment: Instantaneous / InstantaneousVerbose / LinearizedAverage (BMT:505-632) ---------
const SRC18 = (
    :S_phase_change_vap_lcl, :S_phase_change_vap_icl, :S_acnv_lcl_rai, :S_acnv_icl_sno, :S_accr_lcl_rai,
    :S_accr_lcl_sno_cold, :S_accr_lcl_sno_warm, :S_accr_melt_lcl_sno, :S_accr_icl_rai, :S_accr_freeze_icl_rai,
    :S_accr_icl_sno, :S_accr_rai_sno_cold, :S_accr_rai_sno_warm, :S_accr_melt_rai_sno, :S_phase_change_vap_rai,
    :S_phase_change_vap_sno, :S_melt_icl_lcl, :S_melt_sno_rai,
)
const OUT4_1M = (:dq_lcl_dt, :dq_icl_dt, :dq_rai_dt, :dq_sno_dt)

function call_1m(name::Symbol, mp, tps, cols::NTuple{7, Col{FT}}, extra_types, extra_args, n_src) where {FT}
    n = same_length(cols...)
    out = ntuple(_ -> similar(cols[1]), 4)
    src = ntuple(_ -> similar(cols[1]), n_src)
    blk = Ref(pack(FT, mp, tps))
    otab, stab = ptr_table(FT, out), ptr_table(FT, src)
    GC.@preserve blk otab stab begin
        st = if n_src == 0
            ccall((sym(name, FT), libcumicro), Cint,
                (Ptr{Cvoid}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, extra_types...,
                 Ptr{CuPtr{FT}}, Ptr{Cvoid}),
                blk, n, cols..., extra_args..., otab, cur_stream())
        else
            ccall((sym(name, FT), libcumicro), Cint,
                (Ptr{Cvoid}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT},
                 Ptr{CuPtr{FT}}, Ptr{CuPtr{FT}}, Ptr{Cvoid}),
                blk, n, cols..., otab, stab, cur_stream())
        end
    end
    check(st)
    return out, src
end

function BMT.bulk_microphysics_tendencies(
    ::BMT.Instantaneous, ::BMT.Microphysics1Moment, mp::CMP.Microphysics1MParams, tps,
    ρ::Col{FT}, T::Col{FT}, q_tot::Col{FT}, q_lcl::Col{FT}, q_icl::Col{FT}, q_rai::Col{FT}, q_sno::Col{FT},
) where {FT <: FTs}
    out, _ = call_1m(:cumicro_bmt1m_inst, mp, tps, (ρ, T, q_tot, q_lcl, q_icl, q_rai, q_sno), (), (), 0)
    return NamedTuple{OUT4_1M}(out)
end

function BMT.bulk_microphysics_tendencies(
    ::BMT.InstantaneousVerbose, ::BMT.Microphysics1Moment, mp::CMP.Microphysics1MParams, tps,
    ρ::Col{FT}, T::Col{FT}, q_tot::Col{FT}, q_lcl::Col{FT}, q_icl::Col{FT}, q_rai::Col{FT}, q_sno::Col{FT},
) where {FT <: FTs}
    out, src = call_1m(:cumicro_bmt1m_verbose, mp, tps, (ρ, T, q_tot, q_lcl, q_icl, q_rai, q_sno), (), (), 18)
    return merge(NamedTuple{OUT4_1M}(out), NamedTuple{SRC18}(src))
end

function BMT.bulk_microphysics_tendencies(
    ::BMT.LinearizedAverage, ::BMT.Microphysics1Moment, mp::CMP.Microphysics1MParams, tps,
    ρ::Col{FT}, T::Col{FT}, q_tot::Col{FT}, q_lcl::Col{FT}, q_icl::Col{FT}, q_rai::Col{FT}, q_sno::Col{FT},
    Δt::Real, nsub::Integer = 1,
) where {FT <: FTs}
    out, _ = call_1m(:cumicro_bmt1m_linavg, mp, tps, (ρ, T, q_tot, q_lcl, q_icl, q_rai, q_sno), (FT, Cint), (FT(Δt), Cint(nsub)), 0)
    return NamedTuple{OUT4_1M}(out)
end

# ---- 0-moment (BMT:658-680) -------------------------------------------------------------------
function BMT.bulk_microphysics_tendencies(
    ::BMT.Microphysics0Moment, mp::CMP.Microphysics0MParams, tps,
    T::Col{FT}, q_lcl::Col{FT}, q_icl::Col{FT}, q_vap_sat::Union{Col{FT}, Nothing} = nothing,
) where {FT <: FTs}
    n = same_length(T, q_lcl, q_icl, q_vap_sat)
    out = similar(T)
    blk = Ref(pack(FT, mp, tps))
    GC.@preserve blk begin
        st = ccall((sym(:cumicro_bmt0m, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Ptr{Cvoid}),
            blk, n, q_lcl, q_icl, dev(FT, q_vap_sat), out, cur_stream())
    end
    check(st)
    return (; dq_tot_dt = out)
end

# =========================================================================================
# §3.4 stand-alone entry points
# =========================================================================================

# ---- 2-moment terminal velocities (CM2:647-719) -----------------------------------------------
function termvel_2m(name::Symbol, p1, p2, q::Col{FT}, ρ::Col{FT}, N::Col{FT}) where {FT}
    n = same_length(q, ρ, N)
    v0, v1 = similar(q), similar(q)
    r1, r2 = Ref(p1), Ref(p2)
    GC.@preserve r1 r2 begin
        st = ccall((sym(name, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Ptr{Cvoid}),
            r1, r2, n, q, ρ, N, v0, v1, cur_stream())
    end
    check(st)
    return (v0, v1)
end
CM2.rain_terminal_velocity(sb::CMP.SB2006, vel::CMP.SB2006VelType, q_rai::Col{FT}, ρ::Col{FT}, N_rai::Col{FT}) where {FT <: FTs} =
    termvel_2m(:cumicro_termvel_2m_rain_sb, pack_pdf_r(FT, sb.pdf_r), pack_vel(FT, vel), q_rai, ρ, N_rai)
CM2.rain_terminal_velocity(sb::CMP.SB2006, vel::CMP.Chen2022VelTypeRain, q_rai::Col{FT}, ρ::Col{FT}, N_rai::Col{FT}) where {FT <: FTs} =
    termvel_2m(:cumicro_termvel_2m_rain_chen, pack_pdf_r(FT, sb.pdf_r), pack_vel(FT, vel), q_rai, ρ, N_rai)
CM2.cloud_terminal_velocity(pdf_c::CMP.CloudParticlePDF_SB2006, vel::CMP.StokesRegimeVelType, q_lcl::Col{FT}, ρ::Col{FT}, N_lcl::Col{FT}) where {FT <: FTs} =
    termvel_2m(:cumicro_termvel_2m_cloud, pack_pdf_c(FT, pdf_c), pack_vel(FT, vel), q_lcl, ρ, N_lcl)

# ---- the 15 SB2006 process rates in one pass (test/gpu_tests.jl:220-235 calls them one by one) ------------------
const SB2006_LEAVES = (
    :cond_dq_lcl, :evap_dn_rai, :evap_dq_rai, :acnv_dq_lcl, :acnv_dn_lcl, :acnv_dq_rai, :acnv_dn_rai, :lcl_selfcol,
    :accr_dq_lcl, :accr_dn_lcl, :accr_dq_rai, :rai_selfcol, :rai_breakup, :numadj_lcl, :numadj_rai,
)
function sb2006_process_rates(mp::CMP.Microphysics2MParams{WR, Nothing}, tps, ρ::Col{FT}, T::Col{FT}, q_tot::Col{FT},
                              q_lcl::Col{FT}, n_lcl::Col{FT}, q_rai::Col{FT}, n_rai::Col{FT}) where {WR, FT <: FTs}
    n = same_length(ρ, T, q_tot, q_lcl, n_lcl, q_rai, n_rai)
    out = ntuple(_ -> similar(ρ), length(SB2006_LEAVES))
    blk = Ref(pack(FT, mp, tps))
    otab = ptr_table(FT, out)
    GC.@preserve blk otab begin
        st = ccall((sym(:cumicro_sb2006_leaves, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Ptr{CuPtr{FT}}, Ptr{Cvoid}),
            blk, n, ρ, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, otab, cur_stream())
    end
    check(st)
    return NamedTuple{SB2006_LEAVES}(out)
end

# CM2.rain_evaporation / autoconversion / accretion ... over columns: one column of the leaf kernel each.
# (N_* are number densities [1/m³] in the reference's leaf signatures; the kernel takes specific n = N/ρ.)
function leaf_call(mp, tps, ρ::Col{FT}, T, q_tot, q_lcl, N_lcl, q_rai, N_rai, wanted::Tuple) where {FT}
    r = sb2006_process_rates(mp, tps, ρ, T, q_tot, q_lcl, N_lcl ./ ρ, q_rai, N_rai ./ ρ)
    return map(k -> getproperty(r, k), wanted)
end

# ---- CM2.rain_evaporation and its leading-order derivatives over columns (CM2:780-853) ------------------
function rain_evaporation_columns(sb::CMP.SB2006, aps::CMP.AirProperties, tps::TDI.PS, cols::NTuple{8, Col{FT}}) where {FT}
    n = same_length(cols...)
    out = ntuple(_ -> similar(cols[1]), 4)
    # the entry point reads sb, aps and tps of the warm-rain block; the relaxation time scales are not used by this leaf
    blk = Ref(CParams2mWarm{FT}(pack_thermo(FT, tps), pack_sb2006(FT, sb), pack_air(FT, aps), one(FT), one(FT)))
    itab, otab = ptr_table(FT, cols), ptr_table(FT, out)
    GC.@preserve blk itab otab begin
        st = ccall((sym(:cumicro_rain_evaporation_2m, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Int64, Ptr{CuPtr{FT}}, Ptr{CuPtr{FT}}, Ptr{Cvoid}), blk, n, itab, otab, cur_stream())
    end
    check(st)
    return out
end
function CM2.rain_evaporation(sb::CMP.SB2006, aps::CMP.AirProperties, tps::TDI.PS, q_tot::Col{FT}, q_lcl::Col{FT}, q_icl::Col{FT},
                              q_rai::Col{FT}, q_sno::Col{FT}, ρ::Col{FT}, N_rai::Col{FT}, T::Col{FT}) where {FT <: FTs}
    o = rain_evaporation_columns(sb, aps, tps, (q_tot, q_lcl, q_icl, q_rai, q_sno, ρ, N_rai, T))
    return (; ∂ₜρn_rai = o[1], ∂ₜq_rai = o[2])
end
function CM2.∂rain_evaporation_∂N_rai_∂q_rai(sb::CMP.SB2006, aps::CMP.AirProperties, tps::TDI.PS, q_tot::Col{FT}, q_lcl::Col{FT},
                                              q_icl::Col{FT}, q_rai::Col{FT}, q_sno::Col{FT}, ρ::Col{FT}, N_rai::Col{FT}, T::Col{FT}) where {FT <: FTs}
    o = rain_evaporation_columns(sb, aps, tps, (q_tot, q_lcl, q_icl, q_rai, q_sno, ρ, N_rai, T))
    return (; ∂N_rai = o[3], ∂q_rai = o[4])
end

# ---- alternative closures (CM2:920-1002) --------------------------------------------------------
const AltScheme = Union{CMP.KK2000, CMP.B1994, CMP.TC1980, CMP.LD2004}
alt_what_acnv(::CMP.KK2000) = 0
alt_what_acnv(::CMP.B1994) = 1
alt_what_acnv(::CMP.TC1980) = 2
alt_what_acnv(::CMP.LD2004) = 3
alt_what_accr(::CMP.KK2000) = 4
alt_what_accr(::CMP.B1994) = 5
alt_what_accr(::CMP.TC1980) = 6
function alt_call(s, what, smooth, q_lcl::Col{FT}, q_rai, ρ, N_d) where {FT}
    n = same_length(q_lcl, q_rai, ρ, N_d)
    out = similar(q_lcl)
    blk = Ref(pack_alt(FT, s))
    GC.@preserve blk begin
        st = ccall((sym(:cumicro_2m_alt, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Cint, Cint, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Ptr{Cvoid}),
            blk, what, smooth, n, q_lcl, dev(FT, q_rai), dev(FT, ρ), dev(FT, N_d), out, cur_stream())
    end
    check(st)
    return out
end
CM2.conv_q_lcl_to_q_rai(s::AltScheme, q_lcl::Col{FT}, ρ::Col{FT}, N_d::Col{FT}, smooth_transition::Bool = false) where {FT <: FTs} =
    alt_call(s, alt_what_acnv(s), smooth_transition, q_lcl, nothing, ρ, N_d)
CM2.accretion(s::Union{CMP.KK2000, CMP.B1994}, q_lcl::Col{FT}, q_rai::Col{FT}, ρ::Col{FT}) where {FT <: FTs} =
    alt_call(s, alt_what_accr(s), false, q_lcl, q_rai, ρ, nothing)
CM2.accretion(s::CMP.TC1980, q_lcl::Col{FT}, q_rai::Col{FT}) where {FT <: FTs} =
    alt_call(s, alt_what_accr(s), false, q_lcl, q_rai, nothing, nothing)

# ---- 1-moment / non-equilibrium terminal velocities (CM1:240-291, NEQ:250-281) ----------------
function termvel_1m(mp::CMP.Microphysics1MParams, tps, vel, kind::Integer, ρ::Col{FT}, q::Col{FT}) where {FT}
    n = same_length(ρ, q)
    out = similar(q)
    blk = Ref(pack(FT, mp, tps))
    rv = vel === nothing ? nothing : Ref(vel)
    GC.@preserve blk rv begin
        st = ccall((sym(:cumicro_termvel_1m, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Ptr{Cvoid}),
            blk, rv === nothing ? C_NULL : rv, kind, n, ρ, q, out, cur_stream())
    end
    check(st)
    return out
end
# The scalar methods take (precip, vel, ρ, q); the kernel reads the particle parameters from the 1-moment block, so the array
# methods take the full `mp` (and `tps`) in front: CM1.terminal_velocity(mp, tps, mp.precip.rain, vel, ρ, q).
CM1.terminal_velocity(mp::CMP.Microphysics1MParams, tps, ::CMP.Rain, ::CMP.Blk1MVelTypeRain, ρ::Col{FT}, q::Col{FT}) where {FT <: FTs} =
    termvel_1m(mp, tps, nothing, 0, ρ, q)
CM1.terminal_velocity(mp::CMP.Microphysics1MParams, tps, ::CMP.Snow, ::CMP.Blk1MVelTypeSnow, ρ::Col{FT}, q::Col{FT}) where {FT <: FTs} =
    termvel_1m(mp, tps, nothing, 1, ρ, q)
CM1.terminal_velocity(mp::CMP.Microphysics1MParams, tps, ::CMP.Rain, v::CMP.Chen2022VelTypeRain, ρ::Col{FT}, q::Col{FT}) where {FT <: FTs} =
    termvel_1m(mp, tps, pack_vel(FT, v), 2, ρ, q)
CM1.terminal_velocity(mp::CMP.Microphysics1MParams, tps, ::CMP.Snow, v::CMP.Chen2022VelTypeLargeIce, ρ::Col{FT}, q::Col{FT}) where {FT <: FTs} =
    termvel_1m(mp, tps, pack_vel(FT, v), 3, ρ, q)
CMNonEq.terminal_velocity(mp::CMP.Microphysics1MParams, tps, ::CMP.CloudLiquid, v::CMP.StokesRegimeVelType, ρ::Col{FT}, q::Col{FT}) where {FT <: FTs} =
    termvel_1m(mp, tps, pack_vel(FT, v), 4, ρ, q)
CMNonEq.terminal_velocity(mp::CMP.Microphysics1MParams, tps, ::CMP.CloudIce, v::CMP.Chen2022VelTypeSmallIce, ρ::Col{FT}, q::Col{FT}) where {FT <: FTs} =
    termvel_1m(mp, tps, pack_vel(FT, v), 5, ρ, q)

# ---- non-equilibrium condensation / deposition (NEQ:110-224): one source column of the Verbose kernel ----------------
function neq_source(mp::CMP.Microphysics1MParams, tps, micro, thermo, idx::Int)
    ρ = thermo.ρ
    FT = eltype(ρ)
    n = same_length(ρ, thermo.T, micro.q_tot, micro.q_lcl, micro.q_icl, micro.q_rai, micro.q_sno)
    out = similar(ρ)
    src = Any[nothing for _ in 1:18]
    src[idx] = out
    blk = Ref(pack(FT, mp, tps))
    otab, stab = ptr_table(FT, (nothing, nothing, nothing, nothing)), ptr_table(FT, src)
    GC.@preserve blk otab stab begin
        st = ccall((sym(:cumicro_bmt1m_verbose, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT},
             Ptr{CuPtr{FT}}, Ptr{CuPtr{FT}}, Ptr{Cvoid}),
            blk, n, ρ, thermo.T, micro.q_tot, micro.q_lcl, micro.q_icl, micro.q_rai, micro.q_sno, otab, stab, cur_stream())
    end
    check(st)
    return out
end
const ColNT = NamedTuple{<:Any, <:Tuple{Vararg{CuVector}}}
# `mp` carries the selected option and its parameters (mp.processes.cloud_liquid_formation, mp.process_params...)
CMNonEq.conv_q_vap_to_q_lcl(mp::CMP.Microphysics1MParams, tps, micro::ColNT, thermo::ColNT) = neq_source(mp, tps, micro, thermo, 1)
CMNonEq.conv_q_vap_to_q_icl(mp::CMP.Microphysics1MParams, tps, micro::ColNT, thermo::ColNT) = neq_source(mp, tps, micro, thermo, 2)

# ---- ice nucleation leaves (IN:44-584, CO:188-271) ---------------------------------------------
function icenuc_leaf(blk::CParamsIcenuc{FT}, what::Integer, x::Col{FT}, y::Union{Col{FT}, Nothing}, exc = nothing) where {FT}
    n = same_length(x, y)
    out = similar(x)
    r = Ref(blk)
    call(counter) = GC.@preserve r begin
        st = ccall((sym(:cumicro_icenuc, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Cint, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{UInt64}, Ptr{Cvoid}),
            r, what, n, x, dev(FT, y), out, counter, cur_stream())
        check(st)
        out
    end
    return exc === nothing ? call(CuPtr{UInt64}(0)) : with_domain_counter(call, exc)
end
koop_error(n) = DomainError(n, "Δa_w out of range for the Koop 2000 cubic fit at $n point(s)")   # IN:558-562
mohler_error(n) = AssertionError("Si < Sᵢ_max violated at $n point(s)")                          # IN:47, 73

CM_HetIce.deposition_J(dust::CMP.AerosolType, Δa_w::Col{FT}) where {FT <: FTs} =
    icenuc_leaf(pack_icenuc(FT, default_tps(FT); dust), 0, Δa_w, nothing)
CM_HetIce.ABIFM_J(dust::CMP.AerosolType, Δa_w::Col{FT}) where {FT <: FTs} =
    icenuc_leaf(pack_icenuc(FT, default_tps(FT); dust), 1, Δa_w, nothing)
CM_HomIce.homogeneous_J_cubic(koop::CMP.Koop2000, Δa_w::Col{FT}) where {FT <: FTs} =
    icenuc_leaf(pack_icenuc(FT, default_tps(FT); koop), 2, Δa_w, nothing, koop_error)
CM_HomIce.homogeneous_J_linear(koop::CMP.Koop2000, Δa_w::Col{FT}) where {FT <: FTs} =
    icenuc_leaf(pack_icenuc(FT, default_tps(FT); koop), 3, Δa_w, nothing)
CO.a_w_ice(tps::TDI.PS, T::Col{FT}) where {FT <: FTs} = icenuc_leaf(pack_icenuc(FT, tps), 4, T, nothing)
CO.a_w_eT(tps::TDI.PS, e::Col{FT}, T::Col{FT}) where {FT <: FTs} = icenuc_leaf(pack_icenuc(FT, tps), 5, T, e)
CO.a_w_xT(h2so4::CMP.H2SO4SolutionParameters, tps::TDI.PS, x::Col{FT}, T::Col{FT}) where {FT <: FTs} =
    icenuc_leaf(pack_icenuc(FT, tps; h2so4), 6, T, x)
CO.H2SO4_soln_saturation_vapor_pressure(h2so4::CMP.H2SO4SolutionParameters, x::Col{FT}, T::Col{FT}) where {FT <: FTs} =
    icenuc_leaf(pack_icenuc(FT, default_tps(FT); h2so4), 7, T, x)
CM_HetIce.P3_deposition_N_i(mm::CMP.MorrisonMilbrandt2014, T::Col{FT}) where {FT <: FTs} =
    icenuc_leaf(pack_icenuc(FT, default_tps(FT); mm2014 = mm), 8, T, nothing)
CM_HetIce.INP_concentration_mean(f::CMP.Frostenberg2023, T::Col{FT}) where {FT <: FTs} =
    icenuc_leaf(pack_icenuc(FT, default_tps(FT); frostenberg = f), 9, T, nothing)
CM_HetIce.dust_activated_number_fraction(dust::CMP.AerosolType, mohler::CMP.Mohler2006, Si::Col{FT}, T::Col{FT}) where {FT <: FTs} =
    icenuc_leaf(pack_icenuc(FT, default_tps(FT); dust, mohler), 10, Si, T, mohler_error)
# entry points that do not read the thermodynamics block still carry one: any valid parameter set does
default_tps(::Type{FT}) where {FT} = TDP.ThermodynamicsParameters(FT)

function icenuc_rates(blk::CParamsIcenuc{FT}, what::Integer, cols::Tuple, two_outputs::Bool, exc = nothing) where {FT}
    n = same_length(cols...)
    out = similar(cols[1])
    out2 = two_outputs ? similar(cols[1]) : nothing
    r = Ref(blk)
    itab = ptr_table(FT, (cols..., ntuple(_ -> nothing, 5 - length(cols))...))
    call(counter) = GC.@preserve r itab begin
        st = ccall((sym(:cumicro_icenuc_rates, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Cint, Int64, Ptr{CuPtr{FT}}, CuPtr{FT}, CuPtr{FT}, CuPtr{UInt64}, Ptr{Cvoid}),
            r, what, n, itab, out, dev(FT, out2), counter, cur_stream())
        check(st)
        two_outputs ? (out, out2) : out
    end
    return exc === nothing ? call(CuPtr{UInt64}(0)) : with_domain_counter(call, exc)
end
CM_HetIce.MohlerDepositionRate(dust::CMP.AerosolType, mohler::CMP.Mohler2006, Si::Col{FT}, T::Col{FT}, dSi_dt::Col{FT}, N_aer::Col{FT}) where {FT <: FTs} =
    icenuc_rates(pack_icenuc(FT, default_tps(FT); dust, mohler), 0, (Si, T, dSi_dt, N_aer), false, mohler_error)
CM_HetIce.P3_het_N_i(mm::CMP.MorrisonMilbrandt2014, T::Col{FT}, N_l::Col{FT}, V_l::Col{FT}, Δt::Col{FT}) where {FT <: FTs} =
    icenuc_rates(pack_icenuc(FT, default_tps(FT); mm2014 = mm), 1, (T, N_l, V_l, Δt), false)
CM_HetIce.INP_concentration_frequency(f::CMP.Frostenberg2023, INPC::Col{FT}, T::Col{FT}) where {FT <: FTs} =
    icenuc_rates(pack_icenuc(FT, default_tps(FT); frostenberg = f), 2, (INPC, T), false)
function P3.het_ice_nucleation(dust::CMP.AerosolType, tps::TDI.PS, q_lcl::Col{FT}, N_lcl::Col{FT}, RH::Col{FT}, T::Col{FT}, ρₐ::Col{FT}) where {FT <: FTs}
    dNdt, dLdt = icenuc_rates(pack_icenuc(FT, tps; dust), 3, (q_lcl, N_lcl, RH, T, ρₐ), true)
    return (; dNdt, dLdt)
end

# ---- ARG2000 aerosol activation (+ nucleation rates of the same state), AA:138-433 ------------
function arg_icenuc(ap, ad, aps, tps, T::Col{FT}, p::Col{FT}, w::Col{FT}, q_tot::Col{FT}, q_liq::Col{FT}, q_ice::Col{FT},
                    N_liq::Union{Col{FT}, Nothing}, N_ice::Union{Col{FT}, Nothing};
                    want_N = true, want_M = false, want_S = false, dust = nothing, koop = nothing) where {FT}
    n = same_length(T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice)
    nm = AM.n_modes(ad)
    zeros_col() = CUDA.zeros(FT, n)
    N_liq === nothing && (N_liq = zeros_col())      # AA:260-273: the 10-argument methods pass N_liq = N_ice = 0
    N_ice === nothing && (N_ice = zeros_col())
    S = want_S ? similar(T) : nothing
    N_act = want_N ? [similar(T) for _ in 1:nm] : nothing
    M_act = want_M ? [similar(T) for _ in 1:nm] : nothing
    blk = Ref(pack_icenuc(FT, tps; aps, ap, ad, dust, koop))
    ntab = N_act === nothing ? nothing : ptr_table(FT, N_act)
    mtab = M_act === nothing ? nothing : ptr_table(FT, M_act)
    GC.@preserve blk ntab mtab begin
        st = ccall((sym(:cumicro_arg_icenuc, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT},
             CuPtr{FT}, Ptr{CuPtr{FT}}, Ptr{CuPtr{FT}}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{UInt64}, Ptr{Cvoid}),
            blk, n, T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice, dev(FT, S),
            ntab === nothing ? C_NULL : ntab, mtab === nothing ? C_NULL : mtab,
            CuPtr{FT}(0), CuPtr{FT}(0), CuPtr{FT}(0), CuPtr{FT}(0), CuPtr{UInt64}(0), cur_stream())
    end
    check(st)
    return (; S_max = S, N_act, M_act)
end
const AP = CMP.AerosolActivationParameters
const AD = CMP.AerosolDistributionType
AA.max_supersaturation(ap::AP, ad::AD, aps::CMP.AirProperties, tps::TDI.PS, T::Col{FT}, p::Col{FT}, w::Col{FT}, q_tot::Col{FT},
                       q_liq::Col{FT}, q_ice::Col{FT}, N_liq::Union{Col{FT}, Nothing} = nothing, N_ice::Union{Col{FT}, Nothing} = nothing) where {FT <: FTs} =
    arg_icenuc(ap, ad, aps, tps, T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice; want_N = false, want_S = true).S_max
AA.N_activated_per_mode(ap::AP, ad::AD, aps::CMP.AirProperties, tps::TDI.PS, T::Col{FT}, p::Col{FT}, w::Col{FT}, q_tot::Col{FT},
                        q_liq::Col{FT}, q_ice::Col{FT}, N_liq::Union{Col{FT}, Nothing} = nothing, N_ice::Union{Col{FT}, Nothing} = nothing) where {FT <: FTs} =
    Tuple(arg_icenuc(ap, ad, aps, tps, T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice).N_act)
AA.M_activated_per_mode(ap::AP, ad::AD, aps::CMP.AirProperties, tps::TDI.PS, T::Col{FT}, p::Col{FT}, w::Col{FT}, q_tot::Col{FT},
                        q_liq::Col{FT}, q_ice::Col{FT}, N_liq::Union{Col{FT}, Nothing} = nothing, N_ice::Union{Col{FT}, Nothing} = nothing) where {FT <: FTs} =
    Tuple(arg_icenuc(ap, ad, aps, tps, T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice; want_N = false, want_M = true).M_act)
AA.total_N_activated(ap::AP, ad::AD, aps::CMP.AirProperties, tps::TDI.PS, T::Col{FT}, args::Vararg{Union{Col{FT}, Nothing}}) where {FT <: FTs} =
    reduce(+, AA.N_activated_per_mode(ap, ad, aps, tps, T, args...))
AA.total_M_activated(ap::AP, ad::AD, aps::CMP.AirProperties, tps::TDI.PS, T::Col{FT}, args::Vararg{Union{Col{FT}, Nothing}}) where {FT <: FTs} =
    reduce(+, AA.M_activated_per_mode(ap, ad, aps, tps, T, args...))

# ---- the trained-emulator methods of ext/EmulatorModelsExt.jl:32-103 for a multilayer-perceptron machine ------------
"""
    CuMicroEmulatorMLP(layers; activation = :relu, log_features = true, feat_mean, feat_scale, target_transform = false)

The machine of `AA.N_activated_per_mode(machine, ap, ad, aip, tps, T, p, w, qₜ, qₗ, qᵢ)` on the device: dense layers
`layers = [(W₁, b₁), ...]` (`Wₗ` of size outputs × inputs, as `Flux.Dense` stores it) behind the preprocessing of the reference's
training pipeline (`ext/Common.jl:57-77`: log of N, mean and velocity; a standardizer) and its inverse target transform
(`ext/Common.jl:158-160`).  The weights are uploaded once, in the float type of the first call.
"""
struct CuMicroEmulatorMLP
    layers::Vector{Tuple{Matrix{Float64}, Vector{Float64}}}
    activation::Symbol
    log_features::Bool
    feat_mean::Vector{Float64}
    feat_scale::Vector{Float64}
    target_transform::Bool
    device_weights::Dict{DataType, Any}
end
function CuMicroEmulatorMLP(layers; activation = :relu, log_features = true, feat_mean = nothing, feat_scale = nothing, target_transform = false)
    nf = size(layers[1][1], 2)
    CuMicroEmulatorMLP([(Matrix{Float64}(W), Vector{Float64}(b)) for (W, b) in layers], activation, log_features,
        feat_mean === nothing ? zeros(nf) : Vector{Float64}(feat_mean), feat_scale === nothing ? ones(nf) : Vector{Float64}(feat_scale),
        target_transform, Dict{DataType, Any}())
end
# the buffer layout of include/cumicro.h: per layer W as [inputs][outputs] (outputs fastest) then b
packed_weights(::Type{FT}, m::CuMicroEmulatorMLP) where {FT} = FT.(reduce(vcat, [vcat(vec(W), b) for (W, b) in m.layers]))   # vec(W) of an outputs × inputs matrix: outputs fastest
device_weights(::Type{FT}, m::CuMicroEmulatorMLP) where {FT} = get!(() -> CUDA.CuVector{FT}(packed_weights(FT, m)), m.device_weights, FT)
pad_to(::Type{T}, v, n) where {T} = ntuple(i -> i <= length(v) ? T(v[i]) : T(0), n)
function pack_emulator(::Type{FT}, m::CuMicroEmulatorMLP, ap, ad) where {FT}
    nm = AM.n_modes(ad)
    size(m.layers[1][1], 2) == 4nm + 3 || throw(ArgumentError("the emulator was trained for $((size(m.layers[1][1], 2) - 3) ÷ 4) modes, the distribution has $nm"))
    hygro = AA.mean_hygroscopicity_parameter(ap, ad)                                   # EmulatorModelsExt.jl:45
    act = Dict(:relu => 0, :tanh => 1, :logistic => 2, :identity => 3)[m.activation]
    CParamsEmulator{FT}(
        pad_to(FT, [ad.modes[j].N for j in 1:nm], 8), pad_to(FT, [ad.modes[j].r_dry for j in 1:nm], 8),
        pad_to(FT, [ad.modes[j].stdev for j in 1:nm], 8), pad_to(FT, collect(hygro), 8),
        pad_to(FT, m.feat_mean, 35), pad_to(FT, 1 ./ m.feat_scale, 35),
        Int32(nm), Int32(length(m.layers)), pad_to(Int32, [size(W, 1) for (W, _) in m.layers], 4),
        Int32(act), Int32(m.log_features), Int32(m.target_transform))
end
function aa_emulated(m::CuMicroEmulatorMLP, ap, ad, T::Col{FT}, p::Col{FT}, w::Col{FT}; total = false) where {FT}
    n = same_length(T, p, w)
    nm = AM.n_modes(ad)
    N_act = [similar(T) for _ in 1:nm]
    N_tot = total ? similar(T) : nothing
    blk = Ref(pack_emulator(FT, m, ap, ad))
    wts = device_weights(FT, m)
    ntab = ptr_table(FT, N_act)
    GC.@preserve blk ntab begin
        st = ccall((sym(:cumicro_aa_emulated, FT), libcumicro), Cint,
            (Ptr{Cvoid}, CuPtr{FT}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Ptr{CuPtr{FT}}, CuPtr{FT}, Ptr{Cvoid}),
            blk, wts, n, T, p, w, ntab, dev(FT, N_tot), cur_stream())
    end
    check(st)
    return (; N_act, N_tot)
end
# same names and argument order as the reference's extension methods; qₜ, qₗ, qᵢ are unused there as well
AA.N_activated_per_mode(machine::CuMicroEmulatorMLP, ap::AP, ad::AD, aip::CMP.AirProperties, tps::TDI.PS, T::Col{FT}, p::Col{FT}, w::Col{FT},
                        qₜ::Col{FT}, qₗ::Col{FT}, qᵢ::Col{FT}) where {FT <: FTs} = Tuple(aa_emulated(machine, ap, ad, T, p, w).N_act)
AA.total_N_activated(machine::CuMicroEmulatorMLP, ap::AP, ad::AD, aip::CMP.AirProperties, tps::TDI.PS, T::Col{FT}, p::Col{FT}, w::Col{FT},
                     qₜ::Col{FT}, qₗ::Col{FT}, qᵢ::Col{FT}) where {FT <: FTs} = aa_emulated(machine, ap, ad, T, p, w; total = true).N_tot

# ---- fused 1-moment + 2-moment + ice nucleation with domain diagnostics (BASELINE config 5) ------
"""
    fused_tendencies(mp1, mp2, tps, ρ, T, p, w, q_tot, q_lcl, q_icl, q_rai, q_sno, n_lcl, n_rai; ap, ad, dust, koop, hom_linear, diagnostics)

One pass over the 11 state columns: the 1-moment tendencies (BMT:505-514), the 2-moment warm-rain tendencies (BMT:820-854,
q_ice = q_icl + q_sno), J_dep / J_ABIFM / J_hom at Δa_w = a_w_eT - a_w_ice, and (optionally) the slab's domain sums
`diag = (Σρ(dq_rai+dq_sno) 1M, Σρ dq_rai 2M, ΣN_act, n)` as a 4-element device vector: all-reduce it across ranks
(`cumicro_reduce_diagnostics` / `cumicro_nccl_allreduce_f64`, or MPI) for the global diagnostic.
"""
function fused_tendencies(mp1::CMP.Microphysics1MParams, mp2::CMP.Microphysics2MParams{WR, Nothing}, tps,
                          ρ::Col{FT}, T::Col{FT}, p::Col{FT}, w::Col{FT}, q_tot::Col{FT}, q_lcl::Col{FT}, q_icl::Col{FT},
                          q_rai::Col{FT}, q_sno::Col{FT}, n_lcl::Col{FT}, n_rai::Col{FT};
                          aps = mp2.warm_rain.air_properties, ap = nothing, ad = nothing, dust = nothing, koop = nothing,
                          hom_linear = false, diagnostics = true, window = nothing) where {WR, FT <: FTs}
    ins = (ρ, T, p, w, q_tot, q_lcl, q_icl, q_rai, q_sno, n_lcl, n_rai)
    n = same_length(ins...)
    out = ntuple(_ -> similar(ρ), 11)
    window === nothing || diagnostics || throw(ArgumentError("window needs diagnostics = true"))
    diag = diagnostics ? CUDA.zeros(Float64, 4) : nothing
    b1, b2 = Ref(pack(FT, mp1, tps)), Ref(pack(FT, mp2, tps))
    b3 = Ref(pack_icenuc(FT, tps; aps, ap, ad, dust, koop, hom_linear))
    itab, otab = ptr_table(FT, ins), ptr_table(FT, out)
    GC.@preserve b1 b2 b3 itab otab begin
        if window === nothing
            st = ccall((sym(:cumicro_fused_1m2m_icenuc, FT), libcumicro), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{CuPtr{FT}}, Ptr{CuPtr{FT}}, CuPtr{Float64}, Ptr{Cvoid}),
                b1, b2, b3, n, itab, otab, dev(Float64, diag), cur_stream())
        else   # diag = the DOMAIN sums: the exchange is the tail of the call's finish kernel (peer-memory stores over NVLink)
            st = ccall((sym(:cumicro_fused_1m2m_icenuc_p2p, FT), libcumicro), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{CuPtr{FT}}, Ptr{CuPtr{FT}}, CuPtr{Float64}, Ptr{Cvoid}, Ptr{Cvoid}),
                b1, b2, b3, n, itab, otab, diag, window.handle, cur_stream())
        end
    end
    check(st)
    names = (:dq_lcl_dt_1m, :dq_icl_dt_1m, :dq_rai_dt_1m, :dq_sno_dt_1m, :dq_lcl_dt_2m, :dn_lcl_dt_2m, :dq_rai_dt_2m,
             :dn_rai_dt_2m, :J_dep, :J_ABIFM, :J_hom)
    return merge(NamedTuple{names}(out), (; diag))
end

# ---- P3 (src/P3_*.jl) ----------------------------------------------------------------------------
function P3.get_distribution_logλ_from_prognostic(mp::CMP.Microphysics2MParams{WR, ICE}, tps, ρq_ice::Col{FT}, ρn_ice::Col{FT},
                                                  ρq_rim::Col{FT}, ρb_rim::Col{FT}; brent_iters::Integer = 0) where {WR, ICE <: CMP.P3IceParams, FT <: FTs}
    n = same_length(ρq_ice, ρn_ice, ρq_rim, ρb_rim)
    out = similar(ρq_ice)
    blk = Ref(pack(FT, mp, tps))
    GC.@preserve blk begin
        st = ccall((sym(:cumicro_p3_logl, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Cint, CuPtr{FT}, Ptr{Cvoid}),
            blk, n, ρq_ice, ρn_ice, ρq_rim, ρb_rim, brent_iters, out, cur_stream())
    end
    check(st)
    return out
end

# logλ === nothing: the kernel solves get_distribution_logλ_from_prognostic itself (same bits as the stand-alone solve)
function p3_velocities(mp, tps, ρₐ::Col{FT}, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ; quad = mp.ice.quad) where {FT}
    n = same_length(ρₐ, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ)
    v_n, v_m = similar(ρₐ), similar(ρₐ)
    blk = Ref(pack(FT, mp, tps; quad))
    GC.@preserve blk begin
        st = ccall((sym(:cumicro_termvel_p3, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Ptr{Cvoid}),
            blk, n, ρₐ, ρq_ice, ρn_ice, ρq_rim, ρb_rim, dev(FT, logλ), v_n, v_m, cur_stream())
    end
    check(st)
    return (v_n, v_m)
end
const MP3 = CMP.Microphysics2MParams{<:Any, <:CMP.P3IceParams}
P3.ice_terminal_velocity_number_weighted_from_prognostic(mp::MP3, tps, ρₐ::Col{FT}, ρq_ice::Col{FT}, ρn_ice::Col{FT}, ρq_rim::Col{FT},
                                                         ρb_rim::Col{FT}, logλ::Col{FT}; kw...) where {FT <: FTs} =
    p3_velocities(mp, tps, ρₐ, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ; kw...)[1]
P3.ice_terminal_velocity_mass_weighted_from_prognostic(mp::MP3, tps, ρₐ::Col{FT}, ρq_ice::Col{FT}, ρn_ice::Col{FT}, ρq_rim::Col{FT},
                                                       ρb_rim::Col{FT}, logλ::Col{FT}; kw...) where {FT <: FTs} =
    p3_velocities(mp, tps, ρₐ, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ; kw...)[2]

"""
    p3_process_rates(mp, tps, ρ, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, q_rim, b_rim, logλ; quad)

The stand-alone P3 integrals of one state in one launch (BASELINE config 4): bulk velocities (P3_terminal_velocity.jl:73-133),
`ice_melt` (P3_processes.jl:64-94), `ice_self_collection` (:676-712), `bulk_liquid_ice_collision_sources` (:606-655).
`@assert ρw == psd_r.ρw` of :616 is checked on the host (CUMICRO_E_OPTION -> ArgumentError).
"""
function p3_process_rates(mp::MP3, tps, cols::Vararg{Union{Col{FT}, Nothing}, 12}; quad = mp.ice.quad) where {FT <: FTs}   # cols[12] = logλ may be nothing
    n = same_length(cols...)
    out = ntuple(_ -> similar(cols[1]), 12)
    blk = Ref(pack(FT, mp, tps; quad))
    itab, otab = ptr_table(FT, cols), ptr_table(FT, out)
    GC.@preserve blk itab otab begin
        st = ccall((sym(:cumicro_p3_rates, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Int64, Ptr{CuPtr{FT}}, Ptr{CuPtr{FT}}, Ptr{Cvoid}), blk, n, itab, otab, cur_stream())
    end
    check(st)
    names = (:v_n, :v_m, :melt_dNdt, :melt_dLdt, :selfcol_dNdt, :∂ₜq_c, :∂ₜq_r, :∂ₜN_c, :∂ₜN_r, :∂ₜL_rim, :∂ₜL_ice, :∂ₜB_rim)
    return NamedTuple{names}(out)
end

# P3State thresholds and D_m over columns (P3_particle_properties.jl:43-106, P3_integral_properties.jl:56-61)
function p3_state(mp::MP3, tps, L_ice::Col{FT}, N_ice::Col{FT}, L_rim::Col{FT}, B_rim::Col{FT}, logλ::Col{FT}) where {FT <: FTs}
    n = same_length(L_ice, N_ice, L_rim, B_rim, logλ)
    out = ntuple(_ -> similar(L_ice), 7)
    blk = Ref(pack(FT, mp, tps))
    otab = ptr_table(FT, out)
    GC.@preserve blk otab begin
        st = ccall((sym(:cumicro_p3_state, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Ptr{CuPtr{FT}}, Ptr{Cvoid}),
            blk, n, L_ice, N_ice, L_rim, B_rim, logλ, otab, cur_stream())
    end
    check(st)
    return NamedTuple{(:F_rim, :ρ_rim, :ρ_g, :D_th, :D_gr, :D_cr, :D_m)}(out)
end

# Frostenberg-2023 / Bigg rates of the 2-moment + P3 method on their own (IN:274-511, BMT:998-1075)
function icenuc_f23(mp::MP3, tps, cols::Vararg{Col{FT}, 9}; inpc_log_shift::Union{Col{FT}, Nothing} = nothing) where {FT <: FTs}
    n = same_length(cols..., inpc_log_shift)
    out = ntuple(_ -> similar(cols[1]), 7)
    blk = Ref(pack(FT, mp, tps))
    itab, otab = ptr_table(FT, cols), ptr_table(FT, out)
    GC.@preserve blk itab otab begin
        st = ccall((sym(:cumicro_icenuc_f23, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Int64, Ptr{CuPtr{FT}}, CuPtr{FT}, Ptr{CuPtr{FT}}, Ptr{Cvoid}),
            blk, n, itab, dev(FT, inpc_log_shift), otab, cur_stream())
    end
    check(st)
    names = (:rain_∂ₜn_frz, :rain_∂ₜq_frz, :cloud_∂ₜn_frz, :cloud_∂ₜq_frz, :immersion_limit, :dep_∂ₜn, :dep_∂ₜq)
    return NamedTuple{names}(out)
end

# ---- cloud diagnostics (src/CloudDiagnostics.jl:30-187) ------------------------------------------
function diag_2m(sb::CMP.SB2006, q_lcl::Col{FT}, q_rai, N_lcl, N_rai, ρ, want_Z::Bool, want_r::Bool) where {FT}
    n = same_length(q_lcl, q_rai, N_lcl, N_rai, ρ)
    Z = want_Z ? similar(q_lcl) : nothing
    r = want_r ? similar(q_lcl) : nothing
    pc, pr = Ref(pack_pdf_c(FT, sb.pdf_c)), Ref(pack_pdf_r(FT, sb.pdf_r))
    GC.@preserve pc pr begin
        st = ccall((sym(:cumicro_diag_2m, FT), libcumicro), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Ptr{Cvoid}),
            pc, pr, n, q_lcl, q_rai, N_lcl, N_rai, ρ, dev(FT, Z), dev(FT, r), cur_stream())
    end
    check(st)
    return Z, r
end
CMD.radar_reflectivity_2M(sb::CMP.SB2006, q_lcl::Col{FT}, q_rai::Col{FT}, N_lcl::Col{FT}, N_rai::Col{FT}, ρ_air::Col{FT}) where {FT <: FTs} =
    diag_2m(sb, q_lcl, q_rai, N_lcl, N_rai, ρ_air, true, false)[1]
CMD.effective_radius_2M(sb::CMP.SB2006, q_lcl::Col{FT}, q_rai::Col{FT}, N_lcl::Col{FT}, N_rai::Col{FT}, ρ_air::Col{FT}) where {FT <: FTs} =
    diag_2m(sb, q_lcl, q_rai, N_lcl, N_rai, ρ_air, false, true)[2]
function CMD.radar_reflectivity_1M(mp::CMP.Microphysics1MParams, tps, q_rai::Col{FT}, ρ_air::Col{FT}) where {FT <: FTs}
    n = same_length(q_rai, ρ_air)
    Z = similar(q_rai)
    blk = Ref(pack(FT, mp, tps))
    GC.@preserve blk begin
        st = ccall((sym(:cumicro_diag_1m, FT), libcumicro), Cint, (Ptr{Cvoid}, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Ptr{Cvoid}),
            blk, n, q_rai, ρ_air, Z, cur_stream())
    end
    check(st)
    return Z
end
function CMD.effective_radius_Liu_Hallet_97(p::NamedTuple{(:ρw,)}, ρ_air::Col{FT}, q_lcl::Col{FT},
                                            N_lcl::Union{Col{FT}, Nothing} = nothing, q_rai::Union{Col{FT}, Nothing} = nothing,
                                            N_rai::Union{Col{FT}, Nothing} = nothing) where {FT <: FTs}
    n = same_length(ρ_air, q_lcl, N_lcl, q_rai, N_rai)
    r = similar(q_lcl)
    st = ccall((sym(:cumicro_diag_reff_lh97, FT), libcumicro), Cint,
        (FT, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Ptr{Cvoid}),
        FT(p.ρw), n, ρ_air, q_lcl, dev(FT, N_lcl), dev(FT, q_rai), dev(FT, N_rai), r, cur_stream())
    check(st)
    return r
end

# ---- shared numerics the reference tests on the device (test/gpu_tests.jl:1305-1338) --------------
function p3_leaf(what::Integer, x::Col{FT}, y::Col{FT}) where {FT <: FTs}
    n = same_length(x, y)
    out = similar(x)
    st = ccall((sym(:cumicro_p3_leaf, FT), libcumicro), Cint, (Cint, Int64, CuPtr{FT}, CuPtr{FT}, CuPtr{FT}, Ptr{Cvoid}),
        what, n, x, y, out, cur_stream())
    check(st)
    return out
end

# =========================================================================================
# Multi-GPU: domain diagnostics of the column slabs (SURVEY §8e)
# =========================================================================================
"""
    reduce_diagnostics!(diag::CuVector{Float64}; comm = nothing)

Sum the 4-element diagnostic vector of `fused_tendencies` across the ranks of an NCCL communicator (`comm::Ptr{Cvoid}`, an
`ncclComm_t` created by the host model, e.g. through NCCL.jl) on the current stream; with `comm === nothing` it is the identity
(single GPU).  The only collective of the path: 32 bytes per step.
"""
function reduce_diagnostics!(diag::CuVector{Float64}; comm::Union{Ptr{Cvoid}, Nothing} = nothing)
    comm === nothing && return diag
    st = ccall((:cumicro_nccl_allreduce_f64, libcumicro), Cint, (Ptr{Cvoid}, CuPtr{Float64}, Int64, Ptr{Cvoid}),
        comm, diag, length(diag), cur_stream())
    check(st)
    return diag
end

"""
    P2PWindow(rank, nranks)            # rank is 0-based; allocates this rank's window on the current device
    handle(win)::Vector{UInt8}         # 64 bytes to ship to the other ranks (MPI.Allgather, ...)
    connect!(win, handles)             # handles: nranks x 64 bytes in rank order
    reduce_diagnostics!(diag, win)     # in-place sum over the ranks, one single-block kernel, bit-identical on all ranks
    destroy!(win)                      # after the last call has completed on every rank

The same exchange as `reduce_diagnostics!(diag; comm)` without a library: peer-memory stores over NVLink / NVSwitch.
`fused_tendencies(...; window = win)` does it inside its own finish kernel (no second launch).
"""
mutable struct P2PWindow
    handle::Ptr{Cvoid}
    rank::Int
    nranks::Int
    function P2PWindow(rank::Integer, nranks::Integer)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:cumicro_p2p_window_create, libcumicro), Cint, (Cint, Cint, Ptr{Ptr{Cvoid}}), rank, nranks, h))
        return new(h[], rank, nranks)
    end
end
function handle(win::P2PWindow)
    raw = Vector{UInt8}(undef, 64)
    check(ccall((:cumicro_p2p_window_handle, libcumicro), Cint, (Ptr{Cvoid}, Ptr{UInt8}), win.handle, raw))
    return raw
end
function connect!(win::P2PWindow, handles::AbstractVector{UInt8})
    length(handles) == 64 * win.nranks || throw(ArgumentError("expected $(win.nranks) handles of 64 bytes"))
    raw = Vector{UInt8}(handles)
    check(ccall((:cumicro_p2p_window_connect, libcumicro), Cint, (Ptr{Cvoid}, Ptr{UInt8}), win.handle, raw))
    return win
end
function reduce_diagnostics!(diag::CuVector{Float64}, win::P2PWindow)
    check(ccall((:cumicro_p2p_allreduce_f64, libcumicro), Cint, (Ptr{Cvoid}, CuPtr{Float64}, Cint, Ptr{Cvoid}),
        win.handle, diag, length(diag), cur_stream()))
    return diag
end
function destroy!(win::P2PWindow)
    win.handle == C_NULL || ccall((:cumicro_p2p_window_destroy, libcumicro), Cint, (Ptr{Cvoid},), win.handle)
    win.handle = C_NULL
    return nothing
end

end # module
