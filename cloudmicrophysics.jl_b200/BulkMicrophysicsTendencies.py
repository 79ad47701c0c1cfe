"""Array-level mirror of ``CloudMicrophysics.BulkMicrophysicsTendencies`` (``BMT``).

Same names and argument order as the reference's pointwise methods
(src/BulkMicrophysicsTendencies.jl); every state argument is a structure-of-arrays
device column (1-D contiguous CUDA tensor) and the result is a ``Tendencies`` mapping
of device columns.  One call = one fused sm_100a kernel through the C-ABI of
``libcumicro.so`` — this is what the Julia package extension does with ``ccall``
(INTEGRATION.md).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from . import parameters as CMP
from ._columns import Tendencies, check_columns, ptr, ptr_table, stream_handle, zero_column


# --- dispatch singletons (BMT:64-115) ---------------------------------------------------
class MicrophysicsScheme:
    pass


class Microphysics0Moment(MicrophysicsScheme):
    pass


class Microphysics1Moment(MicrophysicsScheme):
    pass


class Microphysics2Moment(MicrophysicsScheme):
    pass


class TendencyMode:
    pass


class Instantaneous(TendencyMode):
    pass


class InstantaneousVerbose(TendencyMode):
    pass


class LinearizedAverage(TendencyMode):
    pass


def _alloc(like, k):
    return [torch.empty_like(like) for _ in range(k)]


def bulk_microphysics_tendencies(*args, **kw):
    """``bulk_microphysics_tendencies([mode,] scheme, mp, tps, columns...)``.

    Dispatches like the reference: an optional leading ``TendencyMode`` (default
    ``Instantaneous()``, BMT:667), then the scheme singleton."""
    if args and isinstance(args[0], TendencyMode):
        mode, args = args[0], args[1:]
    else:
        mode = Instantaneous()
    if not args or not isinstance(args[0], MicrophysicsScheme):
        raise TypeError("expected a MicrophysicsScheme singleton")
    scheme, args = args[0], args[1:]
    if isinstance(scheme, Microphysics2Moment):
        if not isinstance(mode, Instantaneous):
            raise TypeError("the 2-moment scheme only has Instantaneous tendencies (BMT:820,898)")
        mp = args[0]
        if mp.ice is None:
            return _bmt_2m_warm(*args, **kw)
        from . import _p3_bmt
        return _p3_bmt.bmt_2m_p3(*args, **kw)
    if isinstance(scheme, Microphysics1Moment):
        from . import _bmt_1m
        return _bmt_1m.bmt_1m(mode, *args, **kw)
    if isinstance(scheme, Microphysics0Moment):
        from . import _bmt_1m
        return _bmt_1m.bmt_0m(*args, **kw)
    raise TypeError(f"unknown scheme {scheme!r}")


def _bmt_2m_warm(mp, tps, rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice=None, n_ice=None, q_rim=None,
                 b_rim=None, logλ=None, inpc_log_shift=None, w=None, p=None, *, out=None, materialize_zeros=False):
    """BMT:820-854 — 2-moment warm rain (SB2006).  ``n_ice … p`` are accepted and
    ignored exactly as in the reference (they do not enter the warm-only method);
    ``q_ice`` (optional) enters the thermodynamics."""
    names = ["rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai"]
    cols = [rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai]
    if q_ice is not None:
        names.append("q_ice")
        cols.append(q_ice)
    suf, n, dev = check_columns(cols, names)
    block = CMP.pack_2m_warm(mp, tps)
    if not type(block).__name__.endswith(suf):
        raise TypeError(f"parameter float type does not match the columns ({suf})")
    outs = list(out) if out is not None else _alloc(rho, 4)
    check_columns([rho] + outs, ["rho"] + ["out"] * 4)
    zeros = None
    ztab = None
    if materialize_zeros:
        zeros = _alloc(rho, 4)
        ztab = ptr_table(zeros)
    lib = _abi.load()
    fn = getattr(lib, f"cumicro_bmt2m_warm_{suf}")
    with torch.cuda.device(dev):
        st = fn(C.byref(block), C.c_int64(n), *[ptr(c) for c in cols[:7]], ptr(q_ice), *[ptr(o) for o in outs], ztab,
                stream_handle(dev))
    _abi.check(st, "cumicro_bmt2m_warm")
    z = zeros if zeros is not None else [zero_column(rho)] * 4
    return Tendencies(dq_lcl_dt=outs[0], dn_lcl_dt=outs[1], dq_rai_dt=outs[2], dn_rai_dt=outs[3],
                      dq_ice_dt=z[0], dq_rim_dt=z[1], db_rim_dt=z[2], dn_lcl_activation_dt=z[3])


def bulk_microphysics_tendencies_host(scheme, mp, tps, rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, *, out=None,
                                      chunk=0):
    """Host-buffer form of the 2M warm-rain method: numpy arrays or CPU (ideally
    pinned) torch tensors in and out; the library pipelines H2D / kernel / D2H."""
    import numpy as np
    if not isinstance(scheme, Microphysics2Moment) or mp.ice is not None:
        raise TypeError("host-buffer entry point exists for the 2-moment warm-rain method")
    cols = [rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai]

    def hptr(a):
        if isinstance(a, torch.Tensor):
            assert not a.is_cuda and a.is_contiguous() and a.dim() == 1
            return C.c_void_p(a.data_ptr()), a.shape[0], {torch.float64: "f64", torch.float32: "f32"}[a.dtype]
        assert isinstance(a, np.ndarray) and a.flags.c_contiguous and a.ndim == 1
        return a.ctypes.data_as(C.c_void_p), a.shape[0], {"float64": "f64", "float32": "f32"}[a.dtype.name]

    ptrs = [hptr(a) for a in cols]
    suf, n = ptrs[0][2], ptrs[0][1]
    assert all(p[1] == n and p[2] == suf for p in ptrs)
    if out is None:
        if isinstance(rho, torch.Tensor):
            out = [torch.empty_like(rho).pin_memory() if torch.cuda.is_available() else torch.empty_like(rho) for _ in range(4)]
        else:
            out = [np.empty_like(rho) for _ in range(4)]
    optrs = [hptr(a) for a in out]
    assert all(p[1] == n and p[2] == suf for p in optrs)
    block = CMP.pack_2m_warm(mp, tps)
    if not type(block).__name__.endswith(suf):
        raise TypeError(f"parameter float type does not match the columns ({suf})")
    fn = getattr(_abi.load(), f"cumicro_bmt2m_warm_host_{suf}")
    st = fn(C.byref(block), C.c_int64(n), *[p[0] for p in ptrs], *[p[0] for p in optrs], C.c_int64(chunk))
    _abi.check(st, "cumicro_bmt2m_warm_host")
    return Tendencies(dq_lcl_dt=out[0], dn_lcl_dt=out[1], dq_rai_dt=out[2], dn_rai_dt=out[3])
