"""Array-level mirror of ``CloudMicrophysics.CloudDiagnostics`` (src/CloudDiagnostics.jl): radar reflectivity and
effective radius over device columns, same function names and argument order as the scalar methods."""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from . import parameters as CMP
from ._columns import check_columns, ptr, stream_handle


def _diag_2m(sb, q_lcl, q_rai, N_lcl, N_rai, rho, want_Z, want_reff):
    cols = [q_lcl, q_rai, N_lcl, N_rai, rho]
    suf, n, dev = check_columns(cols, ["q_lcl", "q_rai", "N_lcl", "N_rai", "ρ_air"])
    if not type(sb.pdf_c).__name__.endswith(suf):
        raise TypeError(f"parameter float type does not match the columns ({suf})")
    Z = torch.empty_like(rho) if want_Z else None
    reff = torch.empty_like(rho) if want_reff else None
    with torch.cuda.device(dev):
        st = getattr(_abi.load(), f"cumicro_diag_2m_{suf}")(C.byref(sb.pdf_c), C.byref(sb.pdf_r), C.c_int64(n), *[ptr(c) for c in cols],
                                                           ptr(Z), ptr(reff), stream_handle(dev))
    _abi.check(st, "cumicro_diag_2m")
    return Z, reff


def radar_reflectivity_2M(sb, q_lcl, q_rai, N_lcl, N_rai, rho_air):
    """CMD.radar_reflectivity_2M((; pdf_c, pdf_r)::SB2006, q_lcl, q_rai, N_lcl, N_rai, ρ_air) [dBZ] (CloudDiagnostics.jl:60-79)."""
    return _diag_2m(sb, q_lcl, q_rai, N_lcl, N_rai, rho_air, True, False)[0]


def effective_radius_2M(sb, q_lcl, q_rai, N_lcl, N_rai, rho_air):
    """CMD.effective_radius_2M(sb, q_lcl, q_rai, N_lcl, N_rai, ρ_air) [m] (CloudDiagnostics.jl:95-116)."""
    return _diag_2m(sb, q_lcl, q_rai, N_lcl, N_rai, rho_air, False, True)[1]


def radar_reflectivity_and_effective_radius_2M(sb, q_lcl, q_rai, N_lcl, N_rai, rho_air):
    """Both 2-moment diagnostics from one pass over the five columns (they share the size-distribution parameters)."""
    return _diag_2m(sb, q_lcl, q_rai, N_lcl, N_rai, rho_air, True, True)


def radar_reflectivity_1M(mp, tps, q_rai, rho_air):
    """CMD.radar_reflectivity_1M(rain, q, ρ) [dBZ] (CloudDiagnostics.jl:30-42); ``mp`` is the Microphysics1MParams holding ``rain``."""
    suf, n, dev = check_columns([q_rai, rho_air], ["q_rai", "ρ_air"])
    block = CMP.pack_1m(mp, tps)
    if not type(block).__name__.endswith(suf):
        raise TypeError(f"parameter float type does not match the columns ({suf})")
    Z = torch.empty_like(q_rai)
    with torch.cuda.device(dev):
        st = getattr(_abi.load(), f"cumicro_diag_1m_{suf}")(C.byref(block), C.c_int64(n), ptr(q_rai), ptr(rho_air), ptr(Z), stream_handle(dev))
    _abi.check(st, "cumicro_diag_1m")
    return Z


def effective_radius_Liu_Hallet_97(wtr, rho_air, q_lcl, N_lcl=None, q_rai=None, N_rai=None):
    """CMD.effective_radius_Liu_Hallet_97((; ρw), ρ_air, q_lcl[, N_lcl, q_rai, N_rai]) (CloudDiagnostics.jl:132-165); ``wtr`` is
    anything with ``rho_w`` (the CloudLiquid block) or the density itself."""
    rho_w = float(getattr(wtr, "rho_w", wtr))
    if (N_lcl is None) != (q_rai is None) or (N_lcl is None) != (N_rai is None):
        raise TypeError("effective_radius_Liu_Hallet_97 takes (wtr, ρ_air, q_lcl) or (wtr, ρ_air, q_lcl, N_lcl, q_rai, N_rai)")
    cols = [rho_air, q_lcl] + ([N_lcl, q_rai, N_rai] if N_lcl is not None else [])
    suf, n, dev = check_columns(cols, ["ρ_air", "q_lcl", "N_lcl", "q_rai", "N_rai"])
    out = torch.empty_like(q_lcl)
    crho = C.c_double(rho_w) if suf == "f64" else C.c_float(rho_w)
    with torch.cuda.device(dev):
        st = getattr(_abi.load(), f"cumicro_diag_reff_lh97_{suf}")(crho, C.c_int64(n), ptr(rho_air), ptr(q_lcl), ptr(N_lcl), ptr(q_rai),
                                                                  ptr(N_rai), ptr(out), stream_handle(dev))
    _abi.check(st, "cumicro_diag_reff_lh97")
    return out


def effective_radius_const(cloud_params):
    """CMD.effective_radius_const(cloud_params) = cloud_params.r_eff (CloudDiagnostics.jl:174-179)."""
    return cloud_params.r_eff
