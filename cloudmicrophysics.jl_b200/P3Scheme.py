"""Array methods of the reference's ``CloudMicrophysics.P3Scheme`` module (``src/P3.jl`` and
``src/P3_*.jl``) over structure-of-arrays CUDA columns.

Names and argument order follow the reference's pointwise wrappers that hosts broadcast
(``test/gpu_clima_core_test.jl:35-43, 134-138``):

* ``get_distribution_logλ_from_prognostic(params, ρq_ice, ρn_ice, ρq_rim, ρb_rim)``
  (P3_size_distribution.jl:329-334)
* ``ice_terminal_velocity_number_weighted_from_prognostic`` /
  ``ice_terminal_velocity_mass_weighted_from_prognostic(velocity_params, ρₐ, params, ρq_ice, ρn_ice,
  ρq_rim, ρb_rim, logλ; quad)`` (P3_terminal_velocity.jl:135-173)
* ``process_rates(mp, tps, ρ, T, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, q_rim, b_rim, logλ)``: the
  stand-alone P3 integrals of one state in one launch (velocities, ``ice_melt``,
  ``ice_self_collection``, ``bulk_liquid_ice_collision_sources``).

``params`` / ``velocity_params`` are carried by the flattened block, so these methods take the
``Microphysics2MParams`` object ``mp`` (built ``with_ice=True``) and ``tps``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from . import parameters as CMP
from . import parameters_p3 as CMP3
from ._columns import Tendencies, check_columns, ptr, ptr_table, stream_handle

from .parameters_p3 import ChebyshevGauss, GaussLegendre, build_quadrature  # noqa: F401  (re-exported like P3.jl)

RATE_NAMES = ("v_n", "v_m", "melt_dNdt", "melt_dLdt", "self_collection_dNdt", "dq_c", "dq_r", "dN_c", "dN_r", "dL_rim", "dL_ice",
              "dB_rim")


def _block(mp, tps, suf, quad):
    blk = CMP3.pack_p3(mp, tps, quad=quad)
    if not type(blk).__name__.endswith(suf):
        raise TypeError(f"parameter float type does not match the columns ({suf})")
    return blk


def get_distribution_logλ_from_prognostic(mp, tps, ρq_ice, ρn_ice, ρq_rim, ρb_rim, *, brent_iters=0, out=None):
    cols = [ρq_ice, ρn_ice, ρq_rim, ρb_rim]
    suf, n, dev = check_columns(cols, ["ρq_ice", "ρn_ice", "ρq_rim", "ρb_rim"])
    blk = _block(mp, tps, suf, None)
    out = out if out is not None else torch.empty_like(ρq_ice)
    check_columns([ρq_ice, out], ["ρq_ice", "out"])
    with torch.cuda.device(dev):
        st = getattr(_abi.load(), f"cumicro_p3_logl_{suf}")(C.byref(blk), C.c_int64(n), *[ptr(c) for c in cols], C.c_int(brent_iters),
                                                           ptr(out), stream_handle(dev))
    _abi.check(st, "cumicro_p3_logl")
    return out


get_distribution_loglambda_from_prognostic = get_distribution_logλ_from_prognostic


def _termvel(mp, tps, ρₐ, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ, quad, want_n, want_m):
    cols = [ρₐ, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ]     # logλ = None: solved in the kernel
    suf, n, dev = check_columns(cols[:5] + ([logλ] if logλ is not None else []), ["ρₐ", "ρq_ice", "ρn_ice", "ρq_rim", "ρb_rim", "logλ"])
    blk = _block(mp, tps, suf, quad)
    v_n = torch.empty_like(ρₐ) if want_n else None
    v_m = torch.empty_like(ρₐ) if want_m else None
    with torch.cuda.device(dev):
        st = getattr(_abi.load(), f"cumicro_termvel_p3_{suf}")(C.byref(blk), C.c_int64(n), *[ptr(c) for c in cols], ptr(v_n), ptr(v_m),
                                                              stream_handle(dev))
    _abi.check(st, "cumicro_termvel_p3")
    return v_n, v_m


def ice_terminal_velocity_number_weighted_from_prognostic(mp, tps, ρₐ, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ, *, quad=None):
    return _termvel(mp, tps, ρₐ, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ, quad, True, False)[0]


def ice_terminal_velocity_mass_weighted_from_prognostic(mp, tps, ρₐ, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ, *, quad=None):
    return _termvel(mp, tps, ρₐ, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ, quad, False, True)[1]


def ice_terminal_velocities_from_prognostic(mp, tps, ρₐ, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ=None, *, quad=None):
    """Both weighted velocities from one pass over the quadrature nodes."""
    return _termvel(mp, tps, ρₐ, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ, quad, True, True)


def process_rates(mp, tps, ρ, T, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, q_rim, b_rim, logλ=None, *, quad=None, which=None):
    """Stand-alone P3 integrals (BASELINE config 4) -> Tendencies with RATE_NAMES (``which`` selects a subset).
    ``logλ=None``: the kernel solves ``get_distribution_logλ_from_prognostic`` itself, one point per thread, before the integrals."""
    cols = [ρ, T, ρ, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, q_rim, b_rim, logλ]   # slot 2 (q_tot) is not read
    suf, n, dev = check_columns(cols[:11] + ([logλ] if logλ is not None else []),
                                ["ρ", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice", "q_rim", "b_rim", "logλ"])
    blk = _block(mp, tps, suf, quad)
    names = RATE_NAMES if which is None else tuple(which)
    outs = {k: torch.empty_like(ρ) for k in names}
    with torch.cuda.device(dev):
        st = getattr(_abi.load(), f"cumicro_p3_rates_{suf}")(C.byref(blk), C.c_int64(n), ptr_table(cols),
                                                            ptr_table([outs.get(k) for k in RATE_NAMES]), stream_handle(dev))
    _abi.check(st, "cumicro_p3_rates")
    return Tendencies(**outs)


STATE_NAMES = ("F_rim", "ρ_rim", "ρ_g", "D_th", "D_gr", "D_cr", "D_m")
_LEAF = {"gamma_inc_P": 0, "gamma_inc_Q": 1, "gamma_inc_inv": 2, "rime_mass_fraction": 3, "rime_density": 4}


def state_from_prognostic(mp, tps, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ):
    """``P3.state_from_prognostic`` thresholds (P3_particle_properties.jl:43-56, 101-106) and ``P3.D_m(state, logλ)``
    (P3_integral_properties.jl:56-61) over columns -> Tendencies with STATE_NAMES."""
    cols = [ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ]
    suf, n, dev = check_columns(cols, ["ρq_ice", "ρn_ice", "ρq_rim", "ρb_rim", "logλ"])
    blk = _block(mp, tps, suf, None)
    outs = [torch.empty_like(ρq_ice) for _ in STATE_NAMES]
    with torch.cuda.device(dev):
        st = getattr(_abi.load(), f"cumicro_p3_state_{suf}")(C.byref(blk), C.c_int64(n), *[ptr(c) for c in cols], ptr_table(outs),
                                                            stream_handle(dev))
    _abi.check(st, "cumicro_p3_state")
    return Tendencies(zip(STATE_NAMES, outs))


def D_m(mp, tps, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ):
    return state_from_prognostic(mp, tps, ρq_ice, ρn_ice, ρq_rim, ρb_rim, logλ)["D_m"]


def _leaf(what, x, y):
    suf, n, dev = check_columns([x, y], ["x", "y"])
    out = torch.empty_like(x)
    with torch.cuda.device(dev):
        st = getattr(_abi.load(), f"cumicro_p3_leaf_{suf}")(C.c_int(_LEAF[what]), C.c_int64(n), ptr(x), ptr(y), ptr(out), stream_handle(dev))
    _abi.check(st, "cumicro_p3_leaf")
    return out


def gamma_inc(a, x):
    """``UT.gamma_inc(a, x)`` -> (P, Q) (UT:92-144)."""
    return _leaf("gamma_inc_P", a, x), _leaf("gamma_inc_Q", a, x)


def gamma_inc_inv(a, p):
    """``UT.gamma_inc_inv(a, p, 1 - p)`` (UT:205-252)."""
    return _leaf("gamma_inc_inv", a, p)


def rime_mass_fraction(q_rim, q_ice):
    return _leaf("rime_mass_fraction", q_rim, q_ice)


def rime_density(q_rim, b_rim):
    return _leaf("rime_density", q_rim, b_rim)
