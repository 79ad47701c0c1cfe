"""Fused 1-moment + 2-moment warm rain + ice nucleation (+ ARG2000 activation) over one slab
of grid points, with the domain diagnostics reduced inside the kernel (BASELINE config 5), and
the slab partition / diagnostic all-reduce used on a multi-GPU node.

Every grid point is independent (no halo, no vertical coupling), so the global grid is cut
into contiguous slabs, one per GPU / process; the only collective of the whole path is the
sum of ``NDIAG`` Float64 diagnostics (``torch.distributed`` all-reduce: NCCL over NVLink on
GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from . import parameters as CMP
from ._columns import Tendencies, check_columns, ptr_table, stream_handle

NDIAG = 4
IN_NAMES = ("rho", "T", "p", "w", "q_tot", "q_lcl", "q_icl", "q_rai", "q_sno", "n_lcl", "n_rai")
OUT_NAMES = ("m1_dq_lcl_dt", "m1_dq_icl_dt", "m1_dq_rai_dt", "m1_dq_sno_dt", "m2_dq_lcl_dt", "m2_dn_lcl_dt", "m2_dq_rai_dt",
             "m2_dn_rai_dt", "J_dep", "J_ABIFM", "J_hom")
DIAG_NAMES = ("precip_production_1m", "rain_production_2m", "activated_number", "n_points")


def slab_bounds(n_global: int, world_size: int, rank: int):
    """[lo, hi) of rank's contiguous slab: sizes differ by at most one point, slabs tile
    [0, n_global) exactly and in rank order."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(n_global, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_diagnostics(diag: torch.Tensor, group=None, async_op=False):
    """Sum the per-slab diagnostics over all ranks (in place).  No-op without an initialised
    process group (single-GPU run)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    return dist.all_reduce(diag, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def fused_1m2m_icenuc(mp1, mp2, tps, icenuc_block, rho, T, p, w, q_tot, q_lcl, q_icl, q_rai, q_sno, n_lcl, n_rai, *,
                      out=None, diagnostics=True, reduce_group=None, reduce=True, p2p_window=None):
    """One fused kernel over this rank's slab.  Returns a ``Tendencies`` of the 11 output columns
    plus ``diag`` (Float64 device tensor of NDIAG sums, all-reduced over ``torch.distributed``
    ranks when a process group is initialised; ``reduce=False`` leaves the slab's own sums, e.g. for a
    host that reduces through ``cumicro_nccl_allreduce_f64`` with its own communicator).
    ``p2p_window`` (a connected ``collective.P2PWindow``): the cross-GPU sum happens inside the call's own finish
    kernel by peer-memory stores over NVLink (``cumicro_fused_1m2m_icenuc_p2p_*``); ``diag`` then holds the domain sums
    and no library collective is issued."""
    cols = [rho, T, p, w, q_tot, q_lcl, q_icl, q_rai, q_sno, n_lcl, n_rai]
    suf, n, dev = check_columns(cols, list(IN_NAMES))
    b1, b2 = CMP.pack_1m(mp1, tps), CMP.pack_2m_warm(mp2, tps)
    if not (type(b1).__name__.endswith(suf) and type(icenuc_block).__name__.endswith(suf)):
        raise TypeError(f"parameter float type does not match the columns ({suf})")
    outs = list(out) if out is not None else [torch.empty_like(rho) for _ in OUT_NAMES]
    diag = torch.zeros(NDIAG, dtype=torch.float64, device=dev) if diagnostics else None
    if p2p_window is not None and diag is None:
        raise ValueError("p2p_window needs diagnostics=True")
    with torch.cuda.device(dev):
        if p2p_window is not None:
            fn = getattr(_abi.load(), f"cumicro_fused_1m2m_icenuc_p2p_{suf}")
            st = fn(C.byref(b1), C.byref(b2), C.byref(icenuc_block), C.c_int64(n), ptr_table(cols), ptr_table(outs),
                    C.c_void_p(diag.data_ptr()), p2p_window.handle_, stream_handle(dev))
        else:
            fn = getattr(_abi.load(), f"cumicro_fused_1m2m_icenuc_{suf}")
            st = fn(C.byref(b1), C.byref(b2), C.byref(icenuc_block), C.c_int64(n), ptr_table(cols), ptr_table(outs),
                    C.c_void_p(diag.data_ptr()) if diag is not None else None, stream_handle(dev))
    _abi.check(st, "cumicro_fused_1m2m_icenuc")
    if diag is not None and reduce and p2p_window is None:
        all_reduce_diagnostics(diag, reduce_group)
    res = Tendencies(zip(OUT_NAMES, outs))
    res["diag"] = diag
    return res
