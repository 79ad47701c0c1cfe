"""Array-level mirror of ``CloudMicrophysics.HetIceNucleation`` / ``HomIceNucleation`` and the
water activities of ``Common`` (IN:44-253, 557-584; CO:188-271) over device columns."""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from . import parameters as CMP
from ._columns import check_columns, ptr, stream_handle

WHAT = {"deposition_J": 0, "ABIFM_J": 1, "homogeneous_J_cubic": 2, "homogeneous_J_linear": 3, "a_w_ice": 4, "a_w_eT": 5,
        "a_w_xT": 6, "H2SO4_soln_saturation_vapor_pressure": 7, "P3_deposition_N_i": 8, "INP_concentration_mean": 9,
        "dust_activated_number_fraction": 10}


class DomainError(ValueError):
    """The reference throws DomainError / AssertionError for these inputs (IN:558-562, IN:47)."""


def _leaf(what, blk, x, y=None, check_domain=True):
    cols = [x] + ([y] if y is not None else [])
    suf, n, dev = check_columns(cols, ["x", "y"][: len(cols)])
    out = torch.empty_like(x)
    counter = torch.zeros(1, dtype=torch.int64, device=dev)
    fn = getattr(_abi.load(), f"cumicro_icenuc_{suf}")
    with torch.cuda.device(dev):
        st = fn(C.byref(blk), C.c_int(WHAT[what]), C.c_int64(n), ptr(x), ptr(y), ptr(out), ptr(counter), stream_handle(dev))
    _abi.check(st, "cumicro_icenuc")
    if check_domain:
        nerr = int(counter.item())
        if nerr:
            raise DomainError(f"{what}: {nerr} of {n} points are outside the valid range (outputs are NaN there)")
    return out


def deposition_J(dust, tps, Δa_w):
    """IN.deposition_J(dust, Δa_w) (IN:92-102); zero for unsupported aerosol types."""
    return _leaf("deposition_J", CMP.pack_icenuc(tps, dust=dust), Δa_w)


def ABIFM_J(dust, tps, Δa_w):
    """IN.ABIFM_J(dust, Δa_w) (IN:124-134)."""
    return _leaf("ABIFM_J", CMP.pack_icenuc(tps, dust=dust), Δa_w)


def homogeneous_J_cubic(koop, tps, Δa_w, check_domain=True):
    """HomIceNucleation.homogeneous_J_cubic (IN:557-565); raises DomainError like the reference
    unless ``check_domain=False`` (then out-of-range points are NaN)."""
    return _leaf("homogeneous_J_cubic", CMP.pack_icenuc(tps, koop=koop), Δa_w, check_domain=check_domain)


def homogeneous_J_linear(koop, tps, Δa_w):
    """HomIceNucleation.homogeneous_J_linear (IN:581-584)."""
    return _leaf("homogeneous_J_linear", CMP.pack_icenuc(tps, koop=koop), Δa_w)


def a_w_ice(tps, T):
    """CO.a_w_ice (CO:268-271)."""
    return _leaf("a_w_ice", CMP.pack_icenuc(tps), T)


def a_w_eT(tps, e, T):
    """CO.a_w_eT (CO:256-258)."""
    return _leaf("a_w_eT", CMP.pack_icenuc(tps), T, e)


def a_w_xT(h2so4, tps, x, T):
    """CO.a_w_xT (CO:241-245)."""
    return _leaf("a_w_xT", CMP.pack_icenuc(tps, h2so4=h2so4), T, x)


def H2SO4_soln_saturation_vapor_pressure(h2so4, tps, x, T):
    """CO.H2SO4_soln_saturation_vapor_pressure (CO:188-226)."""
    return _leaf("H2SO4_soln_saturation_vapor_pressure", CMP.pack_icenuc(tps, h2so4=h2so4), T, x)


def P3_deposition_N_i(ip, tps, T):
    """IN.P3_deposition_N_i (IN:162-166)."""
    return _leaf("P3_deposition_N_i", CMP.pack_icenuc(tps, mm2014=ip), T)


def INP_concentration_mean(frostenberg, tps, T):
    """IN.INP_concentration_mean (IN:250-253)."""
    return _leaf("INP_concentration_mean", CMP.pack_icenuc(tps, frostenberg=frostenberg), T)


def dust_activated_number_fraction(dust, ip, tps, Si, T, check_domain=True):
    """IN.dust_activated_number_fraction (IN:44-52)."""
    return _leaf("dust_activated_number_fraction", CMP.pack_icenuc(tps, dust=dust, mohler=ip), Si, T, check_domain=check_domain)


# ---- multi-argument rates (cumicro_icenuc_rates_*) ---------------------------------------------------------
RATES = {"MohlerDepositionRate": 0, "P3_het_N_i": 1, "INP_concentration_frequency": 2, "het_ice_nucleation": 3}


def _rates(what, blk, cols, two=False, check_domain=True):
    from ._columns import ptr_table
    suf, n, dev = check_columns(cols, [f"arg{i}" for i in range(len(cols))])
    out = torch.empty_like(cols[0])
    out2 = torch.empty_like(cols[0]) if two else None
    counter = torch.zeros(1, dtype=torch.int64, device=dev)
    fn = getattr(_abi.load(), f"cumicro_icenuc_rates_{suf}")
    with torch.cuda.device(dev):
        st = fn(C.byref(blk), C.c_int(RATES[what]), C.c_int64(n), ptr_table(list(cols) + [None] * (5 - len(cols))), ptr(out), ptr(out2),
                ptr(counter), stream_handle(dev))
    _abi.check(st, "cumicro_icenuc_rates")
    if check_domain:
        nerr = int(counter.item())
        if nerr:
            raise DomainError(f"{what}: {nerr} of {n} points fail the reference's @assert (outputs are NaN there)")
    return (out, out2) if two else out


def MohlerDepositionRate(dust, ip, tps, Si, T, dSi_dt, N_aer, check_domain=True):
    """IN.MohlerDepositionRate(dust, ip, Si, T, dSi_dt, N_aer) (IN:68-77); ``@assert Si < ip.Sᵢ_max`` -> DomainError."""
    return _rates("MohlerDepositionRate", CMP.pack_icenuc(tps, dust=dust, mohler=ip), [Si, T, dSi_dt, N_aer], check_domain=check_domain)


def P3_het_N_i(ip, tps, T, N_l, V_l, Δt):
    """IN.P3_het_N_i(ip, T, Nₗ, Vₗ, Δt) (IN:202-205)."""
    return _rates("P3_het_N_i", CMP.pack_icenuc(tps, mm2014=ip), [T, N_l, V_l, Δt])


def INP_concentration_frequency(frostenberg, tps, INPC, T):
    """IN.INP_concentration_frequency(params, INPC, T) (IN:219-224)."""
    return _rates("INP_concentration_frequency", CMP.pack_icenuc(tps, frostenberg=frostenberg), [INPC, T])


def het_ice_nucleation(aerosol, tps, q_lcl, N_lcl, RH, T, ρₐ):
    """P3.het_ice_nucleation(aerosol, tps, q_lcl, N_lcl, RH, T, ρₐ) (P3_processes.jl:20-45) -> (dNdt, dLdt)."""
    return _rates("het_ice_nucleation", CMP.pack_icenuc(tps, dust=aerosol), [q_lcl, N_lcl, RH, T, ρₐ], two=True)


F23_OUT = ("rain_dn_frz", "rain_dq_frz", "cloud_dn_frz", "cloud_dq_frz", "immersion_limit_dn", "deposition_dn", "deposition_dq")


def f23_and_bigg_rates(mp, tps, rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, inpc_log_shift=None):
    """``IN.liquid_freezing_rate`` (rain and cloud PSD, IN:274-388), ``IN.immersion_limit_rate`` (IN:420-430) and
    ``IN.deposition_rate`` (IN:491-511) over columns, with the arguments BMT:998-1075 passes (``mp`` built ``with_ice=True``)."""
    from . import parameters_p3 as CMP3
    from ._columns import Tendencies, ptr_table
    cols = [rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice]
    suf, n, dev = check_columns(cols + ([inpc_log_shift] if inpc_log_shift is not None else []), ["col"] * 10)
    blk = CMP3.pack_p3(mp, tps)
    if not type(blk).__name__.endswith(suf):
        raise TypeError(f"parameter float type does not match the columns ({suf})")
    outs = [torch.empty_like(rho) for _ in F23_OUT]
    with torch.cuda.device(dev):
        st = getattr(_abi.load(), f"cumicro_icenuc_f23_{suf}")(C.byref(blk), C.c_int64(n), ptr_table(cols), ptr(inpc_log_shift), ptr_table(outs),
                                                              stream_handle(dev))
    _abi.check(st, "cumicro_icenuc_f23")
    return Tendencies(zip(F23_OUT, outs))
