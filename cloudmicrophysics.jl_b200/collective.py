"""Domain diagnostics of the column slabs: the only exchange step of the path.

``reduce_diagnostics`` sums weighted device columns of ONE slab (fixed order, bit-reproducible);
``NcclComm`` wraps the C-ABI communicator helpers (``cumicro_nccl_*``) a host without an NCCL
binding of its own would use, and ``all_reduce`` the in-place sum over its ranks;
``P2PWindow`` is the same sum as peer-memory stores over NVLink inside one kernel
(``cumicro_p2p_*``, csrc/cm_p2p.cuh) — the fused config-5 entry point takes it and reduces in
its own finish kernel."""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from ._columns import check_columns, ptr, ptr_table, stream_handle

_scratch = {}


def reduce_diagnostics(weight, cols, out=None):
    """out[k] = sum_i weight[i] * cols[k][i] (weight may be None = 1) as a Float64 device vector."""
    names = [f"col{k}" for k in range(len(cols))]
    suf, n, dev = check_columns(list(cols) + ([weight] if weight is not None else []), names + (["weight"] if weight is not None else []))
    lib = _abi.load()
    lib.cumicro_reduce_diagnostics_scratch_bytes.restype = C.c_int64
    key = (dev, torch.cuda.current_stream(dev).cuda_stream, len(cols))
    if key not in _scratch:
        nbytes = int(lib.cumicro_reduce_diagnostics_scratch_bytes(len(cols)))
        _scratch[key] = torch.zeros(nbytes, dtype=torch.uint8, device=dev)      # the ticket (first 16 bytes) starts at zero
    sc = _scratch[key]
    if out is None:
        out = torch.empty(len(cols), dtype=torch.float64, device=dev)
    fn = getattr(lib, f"cumicro_reduce_diagnostics_{suf}")
    with torch.cuda.device(dev):
        st = fn(C.c_int64(n), ptr(weight), ptr_table(list(cols)), C.c_int(len(cols)), ptr(out), ptr(sc), C.c_int64(sc.numel()),
                stream_handle(dev))
    _abi.check(st, "cumicro_reduce_diagnostics")
    return out


class NcclComm:
    """ncclComm_t created through the C-ABI: rank 0 calls ``unique_id()`` and ships the 128 bytes to the other ranks by any means."""

    @staticmethod
    def unique_id() -> bytes:
        raw = (C.c_char * 128)()
        _abi.check(_abi.load().cumicro_nccl_unique_id(raw), "cumicro_nccl_unique_id")
        return bytes(raw.raw)

    def __init__(self, nranks: int, unique_id: bytes, rank: int):
        self.handle = C.c_void_p(0)
        raw = (C.c_char * 128).from_buffer_copy(unique_id)
        _abi.check(_abi.load().cumicro_nccl_comm_init_rank(C.byref(self.handle), nranks, raw, rank), "cumicro_nccl_comm_init_rank")

    def all_reduce(self, buf: torch.Tensor):
        """In-place sum of a Float64 device tensor over the ranks, on the current stream."""
        assert buf.is_cuda and buf.dtype == torch.float64 and buf.is_contiguous()
        st = _abi.load().cumicro_nccl_allreduce_f64(self.handle, ptr(buf), C.c_int64(buf.numel()), stream_handle(buf.device))
        _abi.check(st, "cumicro_nccl_allreduce_f64")
        return buf

    def destroy(self):
        if self.handle:
            _abi.load().cumicro_nccl_comm_destroy(self.handle)
            self.handle = C.c_void_p(0)


class P2PWindow:
    """Peer-memory exchange window of this rank (one process per GPU of one NVLink node).

    ``P2PWindow(rank, nranks)`` allocates the local window on the current device; ``handle()`` is
    the 64-byte inter-process handle the host ships to the other ranks; ``connect(handles)`` maps
    the peers' windows (``handles`` in rank order).  ``connect_with_torch_distributed`` does the
    exchange through an initialised ``torch.distributed`` group."""

    HANDLE_BYTES = 64

    def __init__(self, rank: int, nranks: int):
        self.rank, self.nranks = rank, nranks
        self.handle_ = C.c_void_p(0)
        lib = _abi.load()
        _abi.check(lib.cumicro_p2p_window_create(C.c_int(rank), C.c_int(nranks), C.byref(self.handle_)), "cumicro_p2p_window_create")

    def handle(self) -> bytes:
        raw = (C.c_char * self.HANDLE_BYTES)()
        _abi.check(_abi.load().cumicro_p2p_window_handle(self.handle_, raw), "cumicro_p2p_window_handle")
        return bytes(raw.raw)

    def connect(self, handles):
        blob = b"".join(handles)
        if len(blob) != self.HANDLE_BYTES * self.nranks:
            raise ValueError(f"expected {self.nranks} handles of {self.HANDLE_BYTES} bytes")
        raw = (C.c_char * len(blob)).from_buffer_copy(blob)
        _abi.check(_abi.load().cumicro_p2p_window_connect(self.handle_, raw), "cumicro_p2p_window_connect")
        return self

    def connect_with_torch_distributed(self, group=None):
        import torch.distributed as dist
        if self.nranks == 1:
            return self
        box = [None] * self.nranks
        dist.all_gather_object(box, self.handle(), group=group)
        return self.connect(box)

    def set_timeout(self, seconds: float):
        _abi.check(_abi.load().cumicro_p2p_window_set_timeout(self.handle_, C.c_double(seconds)), "cumicro_p2p_window_set_timeout")
        return self

    def status(self):
        """(calls made, call number whose wait timed out or 0)"""
        calls, bad = C.c_int64(0), C.c_int64(0)
        _abi.check(_abi.load().cumicro_p2p_window_status(self.handle_, C.byref(calls), C.byref(bad)), "cumicro_p2p_window_status")
        return calls.value, bad.value

    def all_reduce(self, buf: torch.Tensor):
        """In-place sum of a Float64 device tensor (<= 16 elements) over the ranks, on the current stream."""
        assert buf.is_cuda and buf.dtype == torch.float64 and buf.is_contiguous()
        st = _abi.load().cumicro_p2p_allreduce_f64(self.handle_, ptr(buf), C.c_int(buf.numel()), stream_handle(buf.device))
        _abi.check(st, "cumicro_p2p_allreduce_f64")
        return buf

    def destroy(self):
        """After the last call's work has completed on every rank (synchronise + barrier first)."""
        if self.handle_:
            _abi.load().cumicro_p2p_window_destroy(self.handle_)
            self.handle_ = C.c_void_p(0)
