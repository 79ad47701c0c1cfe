"""1-moment array methods behind ``BMT.bulk_microphysics_tendencies`` (BMT:505-632)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from . import parameters as CMP
from ._columns import Tendencies, check_columns, ptr, ptr_table, stream_handle

OUT_1M = ("dq_lcl_dt", "dq_icl_dt", "dq_rai_dt", "dq_sno_dt")
# field order of _microphysics_source_terms (BMT:206-216)
SRC_1M = ("S_phase_change_vap_lcl", "S_phase_change_vap_icl", "S_acnv_lcl_rai", "S_acnv_icl_sno", "S_accr_lcl_rai",
          "S_accr_lcl_sno_cold", "S_accr_lcl_sno_warm", "S_accr_melt_lcl_sno", "S_accr_icl_rai", "S_accr_freeze_icl_rai",
          "S_accr_icl_sno", "S_accr_rai_sno_cold", "S_accr_rai_sno_warm", "S_accr_melt_rai_sno", "S_phase_change_vap_rai",
          "S_phase_change_vap_sno", "S_melt_icl_lcl", "S_melt_sno_rai")
NAMES = ["rho", "T", "q_tot", "q_lcl", "q_icl", "q_rai", "q_sno"]


def bmt_1m(mode, mp, tps, rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno, Δt=None, nsub=1, *, out=None):
    from .BulkMicrophysicsTendencies import Instantaneous, InstantaneousVerbose, LinearizedAverage
    cols = [rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno]
    suf, n, dev = check_columns(cols, NAMES)
    block = CMP.pack_1m(mp, tps)
    if not type(block).__name__.endswith(suf):
        raise TypeError(f"parameter float type does not match the columns ({suf})")
    outs = list(out) if out is not None else [torch.empty_like(rho) for _ in range(4)]
    check_columns([rho] + outs, ["rho"] + ["out"] * 4)
    lib = _abi.load()
    res = Tendencies(zip(OUT_1M, outs))
    with torch.cuda.device(dev):
        if isinstance(mode, InstantaneousVerbose):
            src = [torch.empty_like(rho) for _ in SRC_1M]
            st = getattr(lib, f"cumicro_bmt1m_verbose_{suf}")(C.byref(block), C.c_int64(n), *[ptr(c) for c in cols],
                                                              ptr_table(outs), ptr_table(src), stream_handle(dev))
            res.update(zip(SRC_1M, src))
        elif isinstance(mode, LinearizedAverage):
            if Δt is None:
                raise TypeError("LinearizedAverage needs Δt (BMT:572-586)")
            cdt = C.c_double(float(Δt)) if suf == "f64" else C.c_float(float(Δt))
            st = getattr(lib, f"cumicro_bmt1m_linavg_{suf}")(C.byref(block), C.c_int64(n), *[ptr(c) for c in cols], cdt,
                                                             C.c_int(int(nsub)), ptr_table(outs), stream_handle(dev))
        elif isinstance(mode, Instantaneous):
            st = getattr(lib, f"cumicro_bmt1m_inst_{suf}")(C.byref(block), C.c_int64(n), *[ptr(c) for c in cols],
                                                           ptr_table(outs), stream_handle(dev))
        else:
            raise TypeError(f"unknown tendency mode {mode!r}")
    _abi.check(st, "cumicro_bmt1m")
    return res


def source_terms_1m(mp, tps, rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno, which):
    """The named subset ``which`` of the 18 source terms of ``BMT._microphysics_source_terms`` (BMT:141-217) and nothing
    else: one launch of the Verbose kernel with every other output column NULL.  Inputs are clamped to >= 0 as at the
    BMT call sites (BMT:146-151)."""
    cols = [rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno]
    suf, n, dev = check_columns(cols, NAMES)
    block = CMP.pack_1m(mp, tps)
    if not type(block).__name__.endswith(suf):
        raise TypeError(f"parameter float type does not match the columns ({suf})")
    unknown = set(which) - set(SRC_1M)
    if unknown:
        raise KeyError(f"unknown source terms {sorted(unknown)}")
    src = [torch.empty_like(rho) if nm in which else None for nm in SRC_1M]
    with torch.cuda.device(dev):
        st = getattr(_abi.load(), f"cumicro_bmt1m_verbose_{suf}")(C.byref(block), C.c_int64(n), *[ptr(c) for c in cols],
                                                                 ptr_table([None] * 4), ptr_table(src), stream_handle(dev))
    _abi.check(st, "cumicro_bmt1m_verbose")
    return Tendencies({nm: t for nm, t in zip(SRC_1M, src) if t is not None})


def bmt_0m(mp, tps, T, q_lcl, q_icl, q_vap_sat=None, *, out=None):
    """BMT:658-680 -> Microphysics0M.remove_precipitation (src/Microphysics0M.jl:35-46):
    ``-max(0, q_lcl + q_icl - threshold)/τ_precip`` with threshold ``qc_0`` or ``S_0 q_vap_sat``, inputs clamped to >= 0.
    ``mp`` is a ``Microphysics0MParams`` (or a bare ``Parameters0M`` block); ``T`` is accepted and not read, as in the reference."""
    cols = [q_lcl, q_icl] + ([q_vap_sat] if q_vap_sat is not None else [])
    suf, n, dev = check_columns(cols, ["q_lcl", "q_icl", "q_vap_sat"])
    blk = getattr(mp, "precip", mp)
    if not type(blk).__name__.endswith(suf):
        raise TypeError(f"parameter float type does not match the columns ({suf})")
    out = out if out is not None else torch.empty_like(q_lcl)
    check_columns([q_lcl, out], ["q_lcl", "out"])
    with torch.cuda.device(dev):
        st = getattr(_abi.load(), f"cumicro_bmt0m_{suf}")(C.byref(blk), C.c_int64(n), ptr(q_lcl), ptr(q_icl), ptr(q_vap_sat), ptr(out),
                                                         stream_handle(dev))
    _abi.check(st, "cumicro_bmt0m")
    return Tendencies(dq_tot_dt=out)
