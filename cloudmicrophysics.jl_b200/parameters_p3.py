"""Host-side mirror of the P3 parameter types: ``CMP.ParametersP3`` and its members
(``src/parameters/MicrophysicsP3.jl:26-319``), ``CMP.P3IceParams``
(``src/parameters/Microphysics2MParams.jl:56-126``), ``CMP.Chen2022VelType``
(``src/parameters/TerminalVelocity.jl``), ``CMP.RainFreezing`` /
``CMP.NIceProxyDepletion`` (``src/parameters/IceNucleation.jl``) and the quadrature
rules of ``src/Quadrature.jl:166-278`` (``ChebyshevGauss``, ``GaussLegendre``,
``build_quadrature``).

Default values: SURVEY.md §A.2 (``docs/src/P3Scheme.md``; checked on the reference's P3
goldens in tests/test_oracle_p3.py).  ``P3_wet_growth_timescale`` and
``P3_constant_slope_parameterization_value`` are not pinned by any reference test; they are
run-time inputs of the C-ABI.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any

import numpy as np

from . import _abi
from . import parameters as P

P.DEFAULTS.update({
    "BF1995_mass_coeff_alpha": 7.38e-11,
    "BF1995_mass_exponent_beta": 1.9,
    "M1996_area_coeff_gamma": 0.2285,
    "M1996_area_exponent_sigma": 1.88,
    "Heymsfield_mu_coeff1": 0.00191,
    "Heymsfield_mu_coeff2": 0.8,
    "Heymsfield_mu_coeff3": 2.0,
    "Heymsfield_mu_cutoff": 6.0,
    "P3_constant_slope_parameterization_value": 0.0,   # unpinned
    "CL1993_local_rime_density_constant_coeff": 51.0,
    "CL1993_local_rime_density_linear_coeff": 114.0,
    "CL1993_local_rime_density_quadratic_coeff": -5.5,
    "P3_wet_growth_timescale": 100.0,                   # unpinned
})

QUAD_MAX_NODES = 128


class AspectRatio:
    def __repr__(self):
        return type(self).__name__ + "()"


class Oblate(AspectRatio):
    pass


class NoAspectRatio(AspectRatio):
    pass


def ParametersP3(FT=np.float64, slope_law="powerlaw", aspect_ratio=None, overrides=None):
    """``CMP.ParametersP3(FT; slope_law = :powerlaw, aspect_ratio = Oblate())`` (MicrophysicsP3.jl:301-319)."""
    if slope_law not in ("constant", "powerlaw"):
        raise AssertionError("slope_law in (:constant, :powerlaw)")
    aspect_ratio = Oblate() if aspect_ratio is None else aspect_ratio
    if not isinstance(aspect_ratio, AspectRatio):
        raise TypeError("aspect_ratio must be Oblate() or NoAspectRatio()")
    td = P._td(FT, overrides)
    F = td.FT
    beta = td["BF1995_mass_exponent_beta"]
    # MassPowerLaw constructor (:38): α_va = p.α_va * 10^(6β_va - 3), in FT arithmetic
    alpha = F(td["BF1995_mass_coeff_alpha"] * F(10) ** (F(6) * beta - F(3)))
    return _abi.struct("p3_scheme", td.suffix)(
        alpha_va=alpha, beta_va=beta,
        gamma=td["M1996_area_coeff_gamma"], sigma=td["M1996_area_exponent_sigma"],
        slope_a=td["Heymsfield_mu_coeff1"], slope_b=td["Heymsfield_mu_coeff2"], slope_c=td["Heymsfield_mu_coeff3"],
        slope_mu_max=td["Heymsfield_mu_cutoff"], slope_mu_const=td["P3_constant_slope_parameterization_value"],
        vent_a=td["SB2006_ventilation_factor_coeff_av"], vent_b=td["SB2006_ventilation_factor_coeff_bv"],
        rim_a=td["CL1993_local_rime_density_constant_coeff"], rim_b=td["CL1993_local_rime_density_linear_coeff"],
        rim_c=td["CL1993_local_rime_density_quadratic_coeff"], rim_rho_ice=td["density_ice_water"],
        tau_wet=td["P3_wet_growth_timescale"], rho_i=td["density_ice_water"], rho_l=td["density_liquid_water"],
        T_freeze=td["temperature_water_freeze"],
        slope_power_law=1 if slope_law == "powerlaw" else 0,
        aspect_oblate=1 if isinstance(aspect_ratio, Oblate) else 0)


# --- src/Quadrature.jl ------------------------------------------------------------------------
def _cospi(x: float) -> float:
    """cospi with exact argument reduction (Julia's cospi is correctly rounded to < 1 ulp)."""
    x = math.fmod(abs(x), 2.0)
    if x > 1.0:
        x = 2.0 - x
    # x in [0, 1]: use the symmetric form that keeps the small angle
    if x <= 0.25:
        return math.cos(math.pi * x)
    if x < 0.75:
        return math.sin(math.pi * (0.5 - x))
    return -math.cos(math.pi * (1.0 - x))


def _fill_quad(FT, n, gauss_legendre, nodes, weights):
    suf = P.suffix(FT)
    q = _abi.struct("quadrature", suf)()
    if not (1 <= n <= QUAD_MAX_NODES):
        raise ValueError(f"quadrature order must be in 1..{QUAD_MAX_NODES}")
    q.n = int(n)
    q.gauss_legendre = int(gauss_legendre)
    for i in range(n):
        q.nodes[i] = nodes[i]
        q.weights[i] = weights[i]
    return q


def GaussLegendre(FT=np.float64, n=None):
    """``GaussLegendre(FT, n)`` / ``GaussLegendre(n)`` (Quadrature.jl:227-236): nodes / weights in
    Float64 (FastGaussQuadrature.gausslegendre; here numpy's leggauss, both exact to ~1 ulp,
    ascending), then converted to FT."""
    if n is None:
        FT, n = np.float64, FT
    F = np.dtype(FT).type
    x, w = np.polynomial.legendre.leggauss(int(n))
    # symmetrise (FastGaussQuadrature returns exactly antisymmetric nodes / symmetric weights)
    x = 0.5 * (x - x[::-1])
    w = 0.5 * (w + w[::-1])
    return _fill_quad(FT, int(n), 1, [F(v) for v in x], [F(v) for v in w])


def ChebyshevGauss(n, FT=np.float64):
    """``ChebyshevGauss(n)`` (Quadrature.jl:166-173).  The block stores, per node,
    y_i = cospi((2i-1)/(2n)) and the total weight sqrt(1-y_i^2) * (pi/n), evaluated in FT exactly
    as ``integrate`` evaluates them per call (Quadrature.jl:74-80)."""
    F = np.dtype(FT).type
    nodes, weights = [], []
    for i in range(1, int(n) + 1):
        y = F(_cospi(float((F(2) * F(i) - F(1)) / F(2 * n))))
        nodes.append(y)
        weights.append(F(np.sqrt(F(1) - y * y) * (F(np.pi) / F(n))))
    return _fill_quad(FT, int(n), 0, nodes, weights)


def build_quadrature(FT, quadrature_order):
    """``Quadrature.build_quadrature(FT, order)`` (Quadrature.jl:270-278)."""
    if int(quadrature_order) in (16, 32, 40, 64):
        return GaussLegendre(FT, int(quadrature_order))
    return ChebyshevGauss(int(quadrature_order), FT)


@dataclass
class Chen2022VelType_:
    """CMP.Chen2022VelType{rain, small_ice, large_ice} (TerminalVelocity.jl)."""
    rain: Any
    small_ice: Any
    large_ice: Any


def Chen2022VelType(FT=np.float64, overrides=None):
    td = P._td(FT, overrides)
    return Chen2022VelType_(rain=P.Chen2022VelTypeRain(td), small_ice=P.Chen2022VelTypeSmallIce(td),
                            large_ice=P.Chen2022VelTypeLargeIce(td))


@dataclass
class NIceProxyDepletion:
    """CMP.NIceProxyDepletion{τ_act} (IceNucleation.jl)."""
    tau_act: float = 300.0


@dataclass
class RainFreezing_:
    het_a: float
    het_B: float


def RainFreezing(FT=np.float64, overrides=None):
    td = P._td(FT, overrides)
    return RainFreezing_(het_a=td["BarklieGokhale1959_a_parameter"], het_B=td["BarklieGokhale1959_B_parameter"])


@dataclass
class P3IceParams_:
    """CMP.P3IceParams (Microphysics2MParams.jl:56-103)."""
    scheme: Any
    terminal_velocity: Chen2022VelType_
    cloud_pdf: Any
    rain_pdf: Any
    ice_nucleation: Any
    rain_freezing: RainFreezing_
    inp_depletion_model: NIceProxyDepletion
    quadrature_order: int
    quad: Any


def P3IceParams(FT=np.float64, is_limited=True, quadrature_order=16, inp_depletion_model=None, overrides=None,
                slope_law="powerlaw", aspect_ratio=None):
    """``P3IceParams(toml_dict; is_limited, quadrature_order, inp_depletion_model)``
    (Microphysics2MParams.jl:105-126)."""
    td = P._td(FT, overrides)
    return P3IceParams_(
        scheme=ParametersP3(td, slope_law=slope_law, aspect_ratio=aspect_ratio),
        terminal_velocity=Chen2022VelType(td),
        cloud_pdf=P.CloudParticlePDF_SB2006(td),
        rain_pdf=P.RainParticlePDF_SB2006(td, is_limited),
        ice_nucleation=P.FrostenbergParameters(td),
        rain_freezing=RainFreezing(td),
        inp_depletion_model=inp_depletion_model if inp_depletion_model is not None else NIceProxyDepletion(300.0),
        quadrature_order=int(quadrature_order),
        quad=build_quadrature(td.FT, int(quadrature_order)))


def pack_p3(mp, tps, quad=None):
    """Flatten ``(mp::Microphysics2MParams{WR, <:P3IceParams}, tps)`` into cumicro_params_p3.
    ``quad`` overrides ``mp.ice.quad`` (the reference passes ``quad`` as a keyword to the
    stand-alone P3 functions)."""
    if mp.ice is None:
        raise TypeError("Microphysics2MParams was built with with_ice = false")
    suf = P.suffix(mp.FT)
    F = np.dtype(mp.FT).type
    ice = mp.ice
    blk = _abi.struct("params_p3", suf)()
    blk.warm = P.pack_2m_warm(mp, tps)
    # BMT:953-954 uses mp.ice.cloud_pdf / rain_pdf for the P3 processes; they are built from the same
    # dictionary as the warm-rain SB2006 members, and the block carries one copy.
    if bytes(ice.cloud_pdf) != bytes(blk.warm.sb.pdf_c) or bytes(ice.rain_pdf) != bytes(blk.warm.sb.pdf_r):
        raise ValueError("mp.ice.cloud_pdf / rain_pdf differ from mp.warm_rain.seifert_beheng.pdf_c / pdf_r")
    blk.scheme = ice.scheme
    blk.vel_rain = ice.terminal_velocity.rain
    blk.vel_small_ice = ice.terminal_velocity.small_ice
    blk.vel_large_ice = ice.terminal_velocity.large_ice
    blk.ice_nucleation = ice.ice_nucleation
    blk.rain_freezing_het_a = F(ice.rain_freezing.het_a)
    blk.rain_freezing_het_B = F(ice.rain_freezing.het_B)
    blk.tau_act = F(ice.inp_depletion_model.tau_act)
    blk.quad = quad if quad is not None else ice.quad
    return blk
