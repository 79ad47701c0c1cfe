"""Host-side mirror of ``CloudMicrophysics.Parameters`` (``CMP``) for the hot path.

In the reference every parameter struct is built from a ClimaParams TOML
dictionary (``CMP/Parameters.jl:61-74``: ``T(FT) = T(CP.create_toml_dict(FT))``)
and carries a few host-side pre-computed constants (gammas, ventilation
coefficients).  ClimaParams is not part of the reference tree, so the default
values live in ``DEFAULTS`` below (keys = ClimaParams names used by the
reference's ``name_map``s; values = SURVEY.md §A.2, checked against the
reference's golden tests in tests/test_oracle_goldens.py).  At run time the
parameter blocks are INPUTS: the Julia extension fills them from the live
``mp``/``tps`` objects, so these defaults only serve this repo's tests/bench.

Each constructor returns the ctypes POD struct declared in
``include/cumicro_params.inc`` (same field names as the reference, ASCII-fied),
in Float64 (``FT=np.float64``) or Float32.  Derived constants are computed in
``FT`` arithmetic exactly where the reference computes them in ``FT``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Optional

import numpy as np

from . import _abi

# ClimaParams defaults (name -> value).  Sources: SURVEY.md §A.2.
DEFAULTS = {
    # --- Thermodynamics.jl parameters
    "temperature_triple_point": 273.16,
    "thermodynamics_temperature_reference": 273.16,
    "pressure_triple_point": 611.657,
    "temperature_water_freeze": 273.15,
    "gas_constant_vapor": 461.5,          # R_v (verified on a_w / non-eq goldens)
    "gas_constant_dry_air": 287.0,        # R_d
    "isobaric_specific_heat_dry_air": 1004.5,
    "isobaric_specific_heat_vapor": 1859.0,
    "isobaric_specific_heat_liquid": 4181.0,
    "isobaric_specific_heat_ice": 2070.0,   # SURVEY A.2 (ClimaParams: 2100 in older releases)
    "latent_heat_vaporization_at_reference": 2.5008e6,
    "latent_heat_sublimation_at_reference": 2.8344e6,
    "specific_humidity_minimum": 1e-10,   # q_min: not pinned by any reference test
    "gravitational_acceleration": 9.81,
    # --- air properties
    "thermal_conductivity_of_air": 2.4e-2,
    "diffusivity_of_water_vapor": 2.26e-5,
    "kinematic_viscosity_of_air": 1.6e-5,
    "density_liquid_water": 1000.0,
    # --- SB2006
    "SB2006_cloud_gamma_distribution_coeff_nu": 1.0,
    "SB2006_cloud_gamma_distribution_coeff_mu": 1.0,
    "SB2006_cloud_droplets_min_mass": 4.2e-15,
    "SB2006_rain_distribution_coeff_nu": -2.0 / 3.0,
    "SB2006_rain_distribution_coeff_mu": 1.0 / 3.0,
    "SB2006_raindrops_min_mass": 2.6e-10,
    "SB2006_raindrops_max_mass": 5e-6,
    "SB2006_raindrops_size_distribution_coeff_N0_min": 2.5e5,
    "SB2006_raindrops_size_distribution_coeff_N0_max": 2e7,
    "SB2006_raindrops_size_distribution_coeff_lambda_min": 1e3,
    "SB2006_raindrops_size_distribution_coeff_lambda_max": 1e4,
    "SB2006_reference_air_density": 1.225,
    "SB2006_collection_kernel_coeff_kcc": 4.44e9,
    "SB2006_collection_kernel_coeff_kcr": 5.25,
    "SB2006_collection_kernel_coeff_krr": 7.12,
    "SB2006_collection_kernel_coeff_kapparr": 60.7,
    "SB2006_autoconversion_correcting_function_coeff_A": 400.0,
    "SB2006_autoconversion_correcting_function_coeff_a": 0.7,
    "SB2006_autoconversion_correcting_function_coeff_b": 3.0,
    "SB2006_accretion_correcting_function_coeff_tau0": 5e-5,
    "SB2006_accretion_correcting_function_coeff_c": 4.0,
    "SB2006_raindrops_self-collection_coeff_d": -5.0,
    "SB2006_raindrops_equilibrium_mean_diameter": 0.9e-3,
    "SB2006_raindrops_breakup_mean_diameter_threshold": 0.35e-3,
    "SB2006_raindrops_breakup_coeff_kbr": 1000.0,
    "SB2006_raindrops_breakup_coeff_kappabr": 2300.0,
    "SB2006_ventilation_factor_coeff_av": 0.78,
    "SB2006_ventilation_factor_coeff_bv": 0.308,
    "SB2006_rain_evaporation_coeff_alpha": 159.0,
    "SB2006_rain_evaporation_coeff_beta": 0.266,
    "Horn2012_number_concentration_adjustment_timescale": 100.0,
    "condensation_evaporation_timescale": 10.0,
    "sublimation_deposition_timescale": 10.0,
    # --- SB2006 rain terminal velocity
    "SB2006_raindrops_terminal_velocity_coeff_aR": 9.65,
    "SB2006_raindrops_terminal_velocity_coeff_bR": 10.3,
    "SB2006_raindrops_terminal_velocity_coeff_cR": 600.0,
    # --- Chen 2022, Table B1 (rain)
    "Chen2022_table_B1_q_coeff": 0.115231,
    "Chen2022_table_B1_a1_coeff": 0.044612,
    "Chen2022_table_B1_a2_coeff": -0.263166,
    "Chen2022_table_B1_a3_coeff": 4.7178,
    "Chen2022_table_B1_a3_pow_coeff": -0.47335,
    "Chen2022_table_B1_b1_coeff": 2.2955,
    "Chen2022_table_B1_b2_coeff": 2.2955,
    "Chen2022_table_B1_b3_coeff": 1.1451,
    "Chen2022_table_B1_b_rho_coeff": 0.038465,
    "Chen2022_table_B1_c1_coeff": 0.0,
    "Chen2022_table_B1_c2_coeff": 0.184325,
    "Chen2022_table_B1_c3_coeff": 0.184325,
}

# CMP/toml/SB2006_limiters.toml:1-11 — the override file the reference's CPU unit
# tests load (test/microphysics2M_tests.jl:26-31); not used by Microphysics2MParams(FT).
SB2006_LIMITERS_OVERRIDE = {
    "SB2006_raindrops_min_mass": 6.54e-11,
    "SB2006_raindrops_size_distribution_coeff_N0_min": 3.5e5,
    "SB2006_raindrops_size_distribution_coeff_N0_max": 2e11,
    "SB2006_raindrops_size_distribution_coeff_lambda_max": 4e4,
}


def suffix(FT) -> str:
    FT = np.dtype(FT)
    if FT == np.float64:
        return "f64"
    if FT == np.float32:
        return "f32"
    raise TypeError(f"unsupported float type {FT}")


class ParamDict:
    """Stand-in for ``ClimaParams.create_toml_dict(FT; override_file)``."""

    def __init__(self, FT=np.float64, overrides: Optional[dict] = None):
        self.FT = np.dtype(FT).type
        self.values = dict(DEFAULTS)
        if overrides:
            unknown = set(overrides) - set(self.values)
            if unknown:
                raise KeyError(f"unknown parameter names: {sorted(unknown)}")
            self.values.update(overrides)

    def __getitem__(self, name):
        return self.FT(self.values[name])

    @property
    def suffix(self):
        return suffix(self.FT)


def _td(arg, overrides=None) -> ParamDict:
    return arg if isinstance(arg, ParamDict) else ParamDict(arg, overrides)


def _gamma(x):
    return type(x)(math.gamma(float(x)))


def _loggamma(x):
    return type(x)(math.lgamma(float(x)))


# --- Thermodynamics.Parameters.ThermodynamicsParameters ----------------------
def ThermodynamicsParameters(FT=np.float64, overrides=None):
    td = _td(FT, overrides)
    return _abi.struct("thermo", td.suffix)(
        T_0=td["thermodynamics_temperature_reference"],
        T_triple=td["temperature_triple_point"],
        press_triple=td["pressure_triple_point"],
        T_freeze=td["temperature_water_freeze"],
        R_v=td["gas_constant_vapor"],
        R_d=td["gas_constant_dry_air"],
        cp_d=td["isobaric_specific_heat_dry_air"],
        cp_v=td["isobaric_specific_heat_vapor"],
        cp_l=td["isobaric_specific_heat_liquid"],
        cp_i=td["isobaric_specific_heat_ice"],
        LH_v0=td["latent_heat_vaporization_at_reference"],
        LH_s0=td["latent_heat_sublimation_at_reference"],
        q_min=td["specific_humidity_minimum"],
        grav=td["gravitational_acceleration"],
    )


# --- CMP/AirProperties.jl:11-31 ------------------------------------------------
def AirProperties(FT=np.float64, overrides=None):
    td = _td(FT, overrides)
    return _abi.struct("air", td.suffix)(
        K_therm=td["thermal_conductivity_of_air"],
        D_vapor=td["diffusivity_of_water_vapor"],
        nu_air=td["kinematic_viscosity_of_air"],
    )


# --- CMP/Microphysics2M.jl:314-672 -----------------------------------------------
def CloudParticlePDF_SB2006(FT=np.float64, overrides=None):
    td = _td(FT, overrides)
    nu_c = td["SB2006_cloud_gamma_distribution_coeff_nu"]
    mu_c = td["SB2006_cloud_gamma_distribution_coeff_mu"]
    return _abi.struct("sb_pdf_c", td.suffix)(
        nu_c=nu_c, mu_c=mu_c,
        xc_min=td["SB2006_cloud_droplets_min_mass"],
        xc_max=td["SB2006_raindrops_min_mass"],
        rho_w=td["density_liquid_water"],
        loggamma_z1=_loggamma((nu_c + 1) / mu_c),
        loggamma_z2=_loggamma((nu_c + 2) / mu_c),
    )


def RainParticlePDF_SB2006(FT=np.float64, is_limited=True, overrides=None):
    td = _td(FT, overrides)
    return _abi.struct("sb_pdf_r", td.suffix)(
        nu_r=td["SB2006_rain_distribution_coeff_nu"],
        mu_r=td["SB2006_rain_distribution_coeff_mu"],
        xr_min=td["SB2006_raindrops_min_mass"],
        xr_max=td["SB2006_raindrops_max_mass"],
        N0_min=td["SB2006_raindrops_size_distribution_coeff_N0_min"],
        N0_max=td["SB2006_raindrops_size_distribution_coeff_N0_max"],
        lam_min=td["SB2006_raindrops_size_distribution_coeff_lambda_min"],
        lam_max=td["SB2006_raindrops_size_distribution_coeff_lambda_max"],
        rho_w=td["density_liquid_water"],
        rho0=td["SB2006_reference_air_density"],
        limited=1 if is_limited else 0,
    )


def SB2006(FT=np.float64, is_limited=True, overrides=None):
    td = _td(FT, overrides)
    F = td.FT
    suf = td.suffix
    av = td["SB2006_ventilation_factor_coeff_av"]
    bv = td["SB2006_ventilation_factor_coeff_bv"]
    beta = td["SB2006_rain_evaporation_coeff_beta"]
    # CMP/Microphysics2M.jl:566-575 (host-side, in FT arithmetic)
    evap = _abi.struct("sb_evap", suf)(
        av=av, bv=bv,
        alpha=td["SB2006_rain_evaporation_coeff_alpha"],
        beta=beta,
        rho0=td["SB2006_reference_air_density"],
        a_vent_1=av / np.cbrt(F(6)),
        b_vent_1=bv * _gamma(F(5) / F(2) + F(3) / F(2) * beta) / F(6) ** (beta / F(2) + F(1) / F(2)),
        a_vent_0_coeff=av * np.cbrt(F(36)),
        b_vent_0_coeff=bv / F(6) ** (beta / F(2) - F(0.5)),
        beta_vent_0=F(-0.5) + F(1.5) * beta,
    )
    return _abi.struct("sb2006", suf)(
        pdf_c=CloudParticlePDF_SB2006(td),
        pdf_r=RainParticlePDF_SB2006(td, is_limited),
        acnv=_abi.struct("sb_acnv", suf)(
            kcc=td["SB2006_collection_kernel_coeff_kcc"],
            x_star=td["SB2006_raindrops_min_mass"],
            rho0=td["SB2006_reference_air_density"],
            A=td["SB2006_autoconversion_correcting_function_coeff_A"],
            a=td["SB2006_autoconversion_correcting_function_coeff_a"],
            b=td["SB2006_autoconversion_correcting_function_coeff_b"],
        ),
        accr=_abi.struct("sb_accr", suf)(
            kcr=td["SB2006_collection_kernel_coeff_kcr"],
            tau0=td["SB2006_accretion_correcting_function_coeff_tau0"],
            rho0=td["SB2006_reference_air_density"],
            c=td["SB2006_accretion_correcting_function_coeff_c"],
        ),
        self=_abi.struct("sb_self", suf)(
            krr=td["SB2006_collection_kernel_coeff_krr"],
            kappa_rr=td["SB2006_collection_kernel_coeff_kapparr"],
            d=td["SB2006_raindrops_self-collection_coeff_d"],
        ),
        brek=_abi.struct("sb_brek", suf)(
            Deq=td["SB2006_raindrops_equilibrium_mean_diameter"],
            Dr_th=td["SB2006_raindrops_breakup_mean_diameter_threshold"],
            kbr=td["SB2006_raindrops_breakup_coeff_kbr"],
            kappa_br=td["SB2006_raindrops_breakup_coeff_kappabr"],
        ),
        evap=evap,
        numadj_tau=td["Horn2012_number_concentration_adjustment_timescale"],
    )


# --- CMP/TerminalVelocity.jl ---------------------------------------------------------
def SB2006VelType(FT=np.float64, overrides=None):
    td = _td(FT, overrides)
    return _abi.struct("vel_sb2006", td.suffix)(
        rho0=td["SB2006_reference_air_density"],
        aR=td["SB2006_raindrops_terminal_velocity_coeff_aR"],
        bR=td["SB2006_raindrops_terminal_velocity_coeff_bR"],
        cR=td["SB2006_raindrops_terminal_velocity_coeff_cR"],
    )


def StokesRegimeVelType(FT=np.float64, overrides=None):
    td = _td(FT, overrides)
    return _abi.struct("vel_stokes", td.suffix)(
        rho_w=td["density_liquid_water"],
        nu_air=td["kinematic_viscosity_of_air"],
        grav=td["gravitational_acceleration"],
    )


def Chen2022VelTypeRain(FT=np.float64, overrides=None):
    td = _td(FT, overrides)
    return _abi.struct("vel_chen_rain", td.suffix)(
        rho0=td["Chen2022_table_B1_q_coeff"],
        a=[td[f"Chen2022_table_B1_a{i}_coeff"] for i in (1, 2, 3)],
        a3_pow=td["Chen2022_table_B1_a3_pow_coeff"],
        b=[td[f"Chen2022_table_B1_b{i}_coeff"] for i in (1, 2, 3)],
        b_rho=td["Chen2022_table_B1_b_rho_coeff"],
        c=[td[f"Chen2022_table_B1_c{i}_coeff"] for i in (1, 2, 3)],
    )


# --- CMP/Microphysics2MParams.jl ---------------------------------------------------
@dataclass
class WarmRainParams2M:
    """CMP.WarmRainParams2M (Microphysics2MParams.jl:14-31)."""
    seifert_beheng: Any
    air_properties: Any
    condevap_tau_relax: float
    subdep_tau_relax: float


@dataclass
class Microphysics2MParams_:
    """CMP.Microphysics2MParams{WR, ICE} (Microphysics2MParams.jl:128-137)."""
    warm_rain: WarmRainParams2M
    ice: Any  # None (warm rain only) or P3IceParams
    FT: Any = np.float64


def Microphysics2MParams(FT=np.float64, with_ice=False, is_limited=True, quadrature_order=16,
                         overrides=None):
    """``CMP.Microphysics2MParams(FT; with_ice, is_limited, quadrature_order)``
    (Microphysics2MParams.jl:151-162)."""
    td = _td(FT, overrides)
    warm = WarmRainParams2M(
        seifert_beheng=SB2006(td, is_limited),
        air_properties=AirProperties(td),
        condevap_tau_relax=td["condensation_evaporation_timescale"],
        subdep_tau_relax=td["sublimation_deposition_timescale"],
    )
    ice = None
    if with_ice:
        from .parameters_p3 import P3IceParams  # noqa: WPS433 (optional family)
        ice = P3IceParams(td, is_limited=is_limited, quadrature_order=quadrature_order)
    return Microphysics2MParams_(warm_rain=warm, ice=ice, FT=td.FT)


def pack_2m_warm(mp: Microphysics2MParams_, tps):
    """Flatten (mp, tps) into the POD block the kernels read — what the Julia
    extension's packer does field by field."""
    suf = suffix(mp.FT)
    if type(tps) is not _abi.struct("thermo", suf):
        raise TypeError("tps float type does not match mp")
    wr = mp.warm_rain
    return _abi.struct("params_2m_warm", suf)(
        tps=tps, sb=wr.seifert_beheng, aps=wr.air_properties,
        condevap_tau_relax=wr.condevap_tau_relax, subdep_tau_relax=wr.subdep_tau_relax,
    )
