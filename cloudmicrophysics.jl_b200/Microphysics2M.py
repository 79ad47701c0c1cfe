"""Array-level mirror of ``CloudMicrophysics.Microphysics2M`` (``CM2``): the SB2006
leaf process rates and the 2-moment terminal velocities over device columns.

The reference exports these as pointwise methods that hosts broadcast
(test/gpu_tests.jl:220-235); here each name takes device columns.  All leaf rates
of one state come out of ONE fused kernel launch (``sb2006_process_rates``); the
per-process functions below are thin selections of its output columns so the
reference's call sites read the same."""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from . import parameters as CMP
from ._columns import Tendencies, check_columns, ptr, ptr_table, stream_handle


def sb2006_process_rates(mp, tps, rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, which=None):
    """All (or the named subset ``which`` of) ``_abi.SB2006_LEAVES`` for each point."""
    cols = [rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai]
    suf, n, dev = check_columns(cols, ["rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai"])
    block = CMP.pack_2m_warm(mp, tps)
    names = _abi.SB2006_LEAVES
    which = list(names) if which is None else list(which)
    outs = [torch.empty_like(rho) if nm in which else None for nm in names]
    lib = _abi.load()
    fn = getattr(lib, f"cumicro_sb2006_leaves_{suf}")
    with torch.cuda.device(dev):
        st = fn(C.byref(block), C.c_int64(n), *[ptr(c) for c in cols], ptr_table(outs), stream_handle(dev))
    _abi.check(st, "cumicro_sb2006_leaves")
    return Tendencies({nm: o for nm, o in zip(names, outs) if o is not None})


def _rain_evaporation(mp, tps, q_tot, q_lcl, q_icl, q_rai, q_sno, rho, N_rai, T, want):
    cols = [q_tot, q_lcl, q_icl, q_rai, q_sno, rho, N_rai, T]
    suf, n, dev = check_columns(cols, ["q_tot", "q_lcl", "q_icl", "q_rai", "q_sno", "rho", "N_rai", "T"])
    block = CMP.pack_2m_warm(mp, tps)
    outs = [torch.empty_like(rho) if w else None for w in want]
    fn = getattr(_abi.load(), f"cumicro_rain_evaporation_2m_{suf}")
    with torch.cuda.device(dev):
        st = fn(C.byref(block), C.c_int64(n), ptr_table(cols), ptr_table(outs), stream_handle(dev))
    _abi.check(st, "cumicro_rain_evaporation_2m")
    return outs


def rain_evaporation(mp, tps, q_tot, q_lcl, q_icl, q_rai, q_sno, rho, N_rai, T):
    """CM2.rain_evaporation(sb, aps, tps, q_tot, q_lcl, q_icl, q_rai, q_sno, ρ, N_rai, T) (CM2:780-828); ``mp`` carries ``sb``
    and ``aps`` (mp.warm_rain).  Returns (∂ₜρn_rai, ∂ₜq_rai) columns."""
    o = _rain_evaporation(mp, tps, q_tot, q_lcl, q_icl, q_rai, q_sno, rho, N_rai, T, (1, 1, 0, 0))
    return Tendencies(dNrho_dt=o[0], dq_dt=o[1])


def d_rain_evaporation_dN_rai_dq_rai(mp, tps, q_tot, q_lcl, q_icl, q_rai, q_sno, rho, N_rai, T):
    """CM2.∂rain_evaporation_∂N_rai_∂q_rai (CM2:844-853): the leading-order derivatives ∂ₜρn_rai / N_rai and ∂ₜq_rai / q_rai."""
    o = _rain_evaporation(mp, tps, q_tot, q_lcl, q_icl, q_rai, q_sno, rho, N_rai, T, (0, 0, 1, 1))
    return Tendencies(dN_rai=o[2], dq_rai=o[3])


globals()["∂rain_evaporation_∂N_rai_∂q_rai"] = d_rain_evaporation_dN_rai_dq_rai   # the reference's own name


def _termvel(fname, pdf, vel, q, rho, N):
    suf, n, dev = check_columns([q, rho, N], ["q", "rho", "N"])
    vt0, vt1 = torch.empty_like(q), torch.empty_like(q)
    fn = getattr(_abi.load(), f"cumicro_{fname}_{suf}")
    with torch.cuda.device(dev):
        st = fn(C.byref(pdf), C.byref(vel), C.c_int64(n), ptr(q), ptr(rho), ptr(N), ptr(vt0), ptr(vt1),
                stream_handle(dev))
    _abi.check(st, f"cumicro_{fname}")
    return vt0, vt1


def rain_terminal_velocity(sb, vel, q_rai, rho, N_rai):
    """CM2.rain_terminal_velocity(SB2006, vel, q_rai, ρ, N_rai) (CM2:685-719): returns
    (number-weighted, mass-weighted) columns; ``vel`` is SB2006VelType or Chen2022VelTypeRain."""
    kind = type(vel).__name__
    if kind.startswith("cumicro_vel_sb2006"):
        return _termvel("termvel_2m_rain_sb", sb.pdf_r, vel, q_rai, rho, N_rai)
    if kind.startswith("cumicro_vel_chen_rain"):
        return _termvel("termvel_2m_rain_chen", sb.pdf_r, vel, q_rai, rho, N_rai)
    raise TypeError(f"unsupported velocity parameterisation {kind}")


def cloud_terminal_velocity(pdf_c, vel, q_liq, rho, N_liq):
    """CM2.cloud_terminal_velocity(pdf_c, ::StokesRegimeVelType, q_liq, ρₐ, N_liq) (CM2:647-664)."""
    return _termvel("termvel_2m_cloud", pdf_c, vel, q_liq, rho, N_liq)


# ---- alternative closures (CM2:920-1002): KK2000, B1994, TC1980, LD2004 ---------------------------------------
_ALT = {"acnv": {"KK2000": 0, "B1994": 1, "TC1980": 2, "LD2004": 3}, "accr": {"KK2000": 4, "B1994": 5, "TC1980": 6}}


def _alt(scheme, what, smooth, q_lcl, q_rai, rho, N_d):
    cols = [c for c in (q_lcl, q_rai, rho, N_d) if c is not None]
    suf, n, dev = check_columns(cols, ["q_lcl", "column 2", "column 3"])
    if not type(scheme.block).__name__.endswith(suf):
        raise TypeError(f"parameter float type does not match the columns ({suf})")
    out = torch.empty_like(q_lcl)
    with torch.cuda.device(dev):
        st = getattr(_abi.load(), f"cumicro_2m_alt_{suf}")(C.byref(scheme.block), C.c_int(what), C.c_int(int(bool(smooth))), C.c_int64(n),
                                                          ptr(q_lcl), ptr(q_rai), ptr(rho), ptr(N_d), ptr(out), stream_handle(dev))
    _abi.check(st, "cumicro_2m_alt")
    return out


def conv_q_lcl_to_q_rai(scheme, q_lcl, rho, N_d, smooth_transition=False):
    """CM2.conv_q_lcl_to_q_rai(::KK2000 | ::B1994 | ::TC1980 | ::LD2004, q_lcl, ρ, N_d[, smooth_transition]) (CM2:920-969)."""
    try:
        what = _ALT["acnv"][scheme.name]
    except (AttributeError, KeyError):
        raise TypeError(f"no conv_q_lcl_to_q_rai method for {scheme!r}") from None
    if what == 0 and smooth_transition:
        raise TypeError("conv_q_lcl_to_q_rai(::KK2000, q_lcl, ρ, N_d) takes no smooth_transition argument (CM2:920)")
    return _alt(scheme, what, smooth_transition, q_lcl, None, rho, N_d)


def accretion(scheme, q_lcl, q_rai, rho=None):
    """CM2.accretion(::KK2000 | ::B1994, q_lcl, q_rai, ρ) and accretion(::TC1980, q_lcl, q_rai) (CM2:985-1002)."""
    try:
        what = _ALT["accr"][scheme.name]
    except (AttributeError, KeyError):
        raise TypeError(f"no accretion method for {scheme!r}") from None
    if (what == 6) != (rho is None):
        raise TypeError("accretion(::TC1980, q_lcl, q_rai) takes no density; KK2000 and B1994 need ρ (CM2:985-1002)")
    return _alt(scheme, what, False, q_lcl, q_rai, rho, None)
