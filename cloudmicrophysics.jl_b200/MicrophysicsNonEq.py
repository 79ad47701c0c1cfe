"""Array-level mirror of ``CloudMicrophysics.MicrophysicsNonEq`` (src/MicrophysicsNonEq.jl): the relaxation-to-equilibrium
cloud condensate tendencies and the monodisperse cloud terminal velocities over device columns.

``micro`` and ``thermo`` are the reference's NamedTuples as mappings / objects of columns:
``micro = dict(q_tot=, q_lcl=, q_icl=, q_rai=, q_sno=)``, ``thermo = dict(ρ=, T=)`` (``rho`` is accepted for ``ρ``).
The tendencies are the ``S_phase_change_vap_lcl`` / ``S_phase_change_vap_icl`` columns of the 1-moment source-term kernel
(the reference's BMT calls exactly these functions, BMT:165-168), so the specific contents are clamped to >= 0 first,
as at that call site; for non-negative inputs the result is that of the scalar method."""
from __future__ import annotations

from . import parameters as CMP
from ._bmt_1m import source_terms_1m
from .Microphysics1M import terminal_velocity as _tv


def _get(obj, *names):
    for nm in names:
        if isinstance(obj, dict) and nm in obj:
            return obj[nm]
        if hasattr(obj, nm):
            return getattr(obj, nm)
    raise KeyError(names[0])


def _state(micro, thermo):
    return (_get(thermo, "ρ", "rho"), _get(thermo, "T"), _get(micro, "q_tot"), _get(micro, "q_lcl"), _get(micro, "q_icl"),
            _get(micro, "q_rai"), _get(micro, "q_sno"))


def _with_option(mp, **options):
    """``mp`` with the given process options (the scalar methods dispatch on ``opt``, not on ``mp.processes``)."""
    import copy
    new = copy.copy(mp)
    new.block = mp.block.copy()
    new.processes = dict(mp.processes)
    for k, v in options.items():
        setattr(new.block.processes, k, 0 if v is None else v.code)
        new.processes[k] = v
    return new


def conv_q_vap_to_q_lcl(opt, mp, tps, micro, thermo):
    """NEQ.conv_q_vap_to_q_lcl(opt::CloudLiquidFormation | nothing, mp, tps, micro, thermo) (NEQ:110-140)."""
    cols = _state(micro, thermo)
    if opt is None:
        return cols[0].new_zeros(cols[0].shape)
    if not isinstance(opt, CMP.CloudLiquidFormation):
        raise TypeError(f"no conv_q_vap_to_q_lcl method for option {opt!r}")
    return source_terms_1m(_with_option(mp, cloud_liquid_formation=opt), tps, *cols, which=["S_phase_change_vap_lcl"]).S_phase_change_vap_lcl


def conv_q_vap_to_q_icl(opt, mp, tps, micro, thermo):
    """NEQ.conv_q_vap_to_q_icl(opt::ConstantTimescale | TemperatureDependent | nothing, …) with the INP limiter (NEQ:161-224)."""
    cols = _state(micro, thermo)
    if opt is None:
        return cols[0].new_zeros(cols[0].shape)
    if not isinstance(opt, (CMP.ConstantTimescale, CMP.TemperatureDependent)):
        raise TypeError(f"no conv_q_vap_to_q_icl method for option {opt!r}")
    return source_terms_1m(_with_option(mp, cloud_ice_formation=opt), tps, *cols, which=["S_phase_change_vap_icl"]).S_phase_change_vap_icl


def terminal_velocity(mp, tps, sediment, vel, rho, q):
    """NEQ.terminal_velocity(sediment::CloudLiquid | CloudIce, vel, ρₐ, q) (NEQ:250-281); ``sediment`` in
    {'cloud_liquid', 'cloud_ice'}."""
    if sediment not in ("cloud_liquid", "cloud_ice"):
        raise ValueError("sediment must be 'cloud_liquid' or 'cloud_ice'")
    return _tv(mp, tps, sediment, vel, rho, q)
