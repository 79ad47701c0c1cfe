"""Host-side multi-GPU logic on CPU: the slab partition tiles the grid exactly, and the
diagnostic all-reduce over a world_size-2 gloo group sums the per-slab partial sums
(this is the only collective of the path; the GPU runs use the same code over NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_bounds_tile_the_grid(built):
    from cumicro.fused import slab_bounds
    for n in (0, 1, 7, 2 ** 24, 2 ** 28 + 5):
        for world in (1, 2, 3, 4, 8):
            edges = [slab_bounds(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slab_bounds(10, 2, 2)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import cumicro  # noqa: F401
    from cumicro.fused import NDIAG, all_reduce_diagnostics, slab_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 100003
    rng = np.random.default_rng(0)
    field = rng.random((n, NDIAG))                   # per-point contributions, identical on every rank
    lo, hi = slab_bounds(n, world, rank)
    diag = torch.from_numpy(field[lo:hi].sum(axis=0))  # what the kernel epilogue produces for this slab
    all_reduce_diagnostics(diag)
    q.put((rank, diag.numpy().tolist(), field.sum(axis=0).tolist()))
    dist.destroy_process_group()


def test_diagnostic_all_reduce_world_size_2_gloo(built):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, diag, total in got:
        np.testing.assert_allclose(diag, total, rtol=1e-13)
    assert got[0][1] == got[1][1]                      # every rank holds the same global sums


def test_all_reduce_is_a_noop_without_a_process_group(built):
    import torch
    from cumicro.fused import all_reduce_diagnostics
    d = torch.ones(4, dtype=torch.float64)
    assert all_reduce_diagnostics(d) is None and float(d.sum()) == 4.0
