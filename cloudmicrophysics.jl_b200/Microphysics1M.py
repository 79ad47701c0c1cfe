"""Array-level mirror of the stand-alone entry points of ``CloudMicrophysics.Microphysics1M``
(``CM1``) and ``MicrophysicsNonEq``: terminal velocities over device columns."""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from . import parameters as CMP
from ._columns import check_columns, ptr, stream_handle

_KIND = {"rain_blk1m": 0, "snow_blk1m": 1, "rain_chen": 2, "snow_chen": 3, "cloud_liquid_stokes": 4, "cloud_ice_chen": 5}


def _termvel(mp, tps, kind, vel, rho, q):
    suf, n, dev = check_columns([rho, q], ["rho", "q"])
    block = CMP.pack_1m(mp, tps)
    out = torch.empty_like(rho)
    fn = getattr(_abi.load(), f"cumicro_termvel_1m_{suf}")
    with torch.cuda.device(dev):
        st = fn(C.byref(block), C.byref(vel) if vel is not None else None, C.c_int(_KIND[kind]), C.c_int64(n), ptr(rho), ptr(q),
                ptr(out), stream_handle(dev))
    _abi.check(st, "cumicro_termvel_1m")
    return out


def terminal_velocity(mp, tps, species, vel, rho, q):
    """``CM1.terminal_velocity(precip, vel, ρ, q)`` (CM1:240-291) / ``CMNonEq.terminal_velocity(sediment,
    vel, ρₐ, q)`` (NEQ:250-281).  ``species`` in {'rain','snow','cloud_liquid','cloud_ice'};
    ``vel`` = None (Blk1M, from ``mp``) or a Chen2022 / Stokes parameter block."""
    name = type(vel).__name__ if vel is not None else ""
    if species == "rain":
        kind = "rain_blk1m" if vel is None else "rain_chen"
        if vel is not None and not name.startswith("cumicro_vel_chen_rain"):
            raise TypeError("rain: vel must be None (Blk1M) or Chen2022VelTypeRain")
    elif species == "snow":
        kind = "snow_blk1m" if vel is None else "snow_chen"
        if vel is not None and not name.startswith("cumicro_vel_chen_large_ice"):
            raise TypeError("snow: vel must be None (Blk1M) or Chen2022VelTypeLargeIce")
    elif species == "cloud_liquid":
        kind = "cloud_liquid_stokes"
        if not name.startswith("cumicro_vel_stokes"):
            raise TypeError("cloud_liquid: vel must be StokesRegimeVelType")
    elif species == "cloud_ice":
        kind = "cloud_ice_chen"
        if not name.startswith("cumicro_vel_chen_small_ice"):
            raise TypeError("cloud_ice: vel must be Chen2022VelTypeSmallIce")
    else:
        raise ValueError(f"unknown species {species!r}")
    return _termvel(mp, tps, kind, vel, rho, q)
