"""``bulk_microphysics_tendencies(::Microphysics2Moment, mp{WR,<:P3IceParams}, tps, ...)`` (BMT:898-1083)
over CUDA columns."""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from . import parameters_p3 as CMP3
from ._columns import Tendencies, check_columns, ptr, ptr_table, stream_handle, zero_column

OUT = ("dq_lcl_dt", "dn_lcl_dt", "dq_rai_dt", "dn_rai_dt", "dq_ice_dt", "dn_ice_dt", "dq_rim_dt", "db_rim_dt")


def bmt_2m_p3(mp, tps, rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, q_rim, b_rim, logλ, inpc_log_shift=None, w=None,
              p=None, *, out=None, quad=None):
    """``w`` and ``p`` only feed the aerosol activation that the reference leaves at zero (BMT:729, 1077-1078).
    ``logλ=None``: the kernel solves ``P3.get_distribution_logλ_from_prognostic`` itself (same bits as the stand-alone solve)."""
    names = ["rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice", "q_rim", "b_rim", "logλ"]
    cols = [rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, q_rim, b_rim, logλ]
    given = [(c, nm) for c, nm in zip(cols + [inpc_log_shift], names + ["inpc_log_shift"]) if c is not None]
    suf, n, dev = check_columns([c for c, _ in given], [nm for _, nm in given])
    blk = CMP3.pack_p3(mp, tps, quad=quad)
    if not type(blk).__name__.endswith(suf):
        raise TypeError(f"parameter float type does not match the columns ({suf})")
    outs = list(out) if out is not None else [torch.empty_like(rho) for _ in range(8)]
    check_columns([rho] + outs, ["rho"] + ["out"] * 8)
    with torch.cuda.device(dev):
        st = getattr(_abi.load(), f"cumicro_bmt2m_p3_{suf}")(C.byref(blk), C.c_int64(n), ptr_table(cols), ptr(inpc_log_shift),
                                                            ptr_table(outs + [None]), stream_handle(dev))
    _abi.check(st, "cumicro_bmt2m_p3")
    return Tendencies(**dict(zip(OUT, outs)), dn_lcl_activation_dt=zero_column(rho))
