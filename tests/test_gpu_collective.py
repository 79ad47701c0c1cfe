"""The exchange step of the path through the C-ABI: one slab's diagnostic sums (cumicro_reduce_diagnostics_*) and
their NCCL all-reduce over a communicator created with cumicro_nccl_* (2 ranks; skipped on a single-GPU box)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reduce_diagnostics_matches_float64_sums_and_is_reproducible(built, cuda):
    import torch
    from cumicro import collective
    rng = np.random.default_rng(3)
    for n in (1, 1000, (1 << 20) + 7):
        for dtype in (np.float64, np.float32):
            rho = rng.uniform(0.2, 1.3, n).astype(dtype)
            cols = [rng.normal(size=n).astype(dtype) * 1e-6 for _ in range(3)]
            d = [torch.from_numpy(c).to(cuda) for c in cols]
            w = torch.from_numpy(rho).to(cuda)
            a = collective.reduce_diagnostics(w, d).cpu().numpy()
            b = collective.reduce_diagnostics(w, d).cpu().numpy()
            assert np.array_equal(a, b)                                   # fixed reduction order
            ref = [np.sum(rho.astype(np.float64) * c.astype(np.float64)) for c in cols]
            scale = [np.sum(np.abs(rho.astype(np.float64) * c.astype(np.float64))) for c in cols]
            for x, r, s in zip(a, ref, scale):
                assert abs(x - r) <= 1e-13 * s
            u = collective.reduce_diagnostics(None, d[:1]).cpu().numpy()
            assert abs(u[0] - np.sum(cols[0].astype(np.float64))) <= 1e-13 * np.sum(np.abs(cols[0].astype(np.float64)))


def test_fused_diagnostics_equal_the_reduction_of_the_output_columns(built, cuda):
    """The in-kernel sums of the config-5 kernel (diag[0], diag[1]) equal cumicro_reduce_diagnostics over its own output columns."""
    import torch
    from cumicro import collective, fused
    from cumicro.testing import arg_test_distribution, synthetic_states_fused
    CMP = built.CMP
    n = (1 << 18) + 11
    st = synthetic_states_fused(n, seed=9)
    d = {k: torch.from_numpy(v).to(cuda) for k, v in st.items()}
    tps = CMP.ThermodynamicsParameters(np.float64)
    blk3 = CMP.pack_icenuc(tps, ad=arg_test_distribution("kappa"), dust=CMP.DustType("Kaolinite"), hom_linear=True)
    out = fused.fused_1m2m_icenuc(CMP.Microphysics1MParams(np.float64), CMP.Microphysics2MParams(np.float64), tps, blk3,
                                  *[d[k] for k in fused.IN_NAMES])
    red = collective.reduce_diagnostics(d["rho"], [out["m1_dq_rai_dt"], out["m1_dq_sno_dt"], out["m2_dq_rai_dt"]]).cpu().numpy()
    diag = out["diag"].cpu().numpy()
    mag = float((d["rho"] * (out["m1_dq_rai_dt"].abs() + out["m1_dq_sno_dt"].abs())).sum())
    assert abs(diag[0] - (red[0] + red[1])) <= 1e-12 * mag
    assert abs(diag[1] - red[2]) <= 1e-12 * float((d["rho"] * out["m2_dq_rai_dt"].abs()).sum())
    assert diag[3] == n


def test_downdraft_cells_do_not_poison_the_activation_diagnostic(built, cuda):
    """w <= 0 (the reference throws a DomainError in AA.max_supersaturation) must not turn the domain sum into NaN."""
    import torch
    from cumicro import fused
    from cumicro.testing import arg_test_distribution, synthetic_states_fused
    CMP = built.CMP
    n = 1 << 14
    st = synthetic_states_fused(n, seed=2)
    st["w"][::3] *= -1.0
    st["w"][1::7] = 0.0
    d = {k: torch.from_numpy(v).to(cuda) for k, v in st.items()}
    tps = CMP.ThermodynamicsParameters(np.float64)
    blk3 = CMP.pack_icenuc(tps, ad=arg_test_distribution("kappa"), dust=CMP.DustType("Kaolinite"), hom_linear=True)
    mp1, mp2 = CMP.Microphysics1MParams(np.float64), CMP.Microphysics2MParams(np.float64)
    out = fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *[d[k] for k in fused.IN_NAMES])
    diag = out["diag"].cpu().numpy()
    assert np.all(np.isfinite(diag)) and diag[2] > 0
    up = st["w"] > 0
    sub = {k: torch.from_numpy(np.ascontiguousarray(v[up])).to(cuda) for k, v in st.items()}
    ref = fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *[sub[k] for k in fused.IN_NAMES])["diag"].cpu().numpy()
    assert abs(diag[2] - ref[2]) <= 1e-12 * ref[2]          # only the updraft cells activate


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import cumicro  # noqa: F401
    from cumicro import collective
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # only to ship the 128-byte id
    box = [collective.NcclComm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, 0)
    comm = collective.NcclComm(world, box[0], rank)
    buf = torch.tensor([1.0 + rank, 10.0 * (rank + 1), 0.5, float(rank)], dtype=torch.float64, device=f"cuda:{rank}")
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):                                      # a side stream, as the host model would use
        comm.all_reduce(buf)
    side.synchronize()
    q.put((rank, buf.cpu().numpy().tolist()))
    comm.destroy()
    dist.destroy_process_group()


def test_nccl_allreduce_through_the_c_abi_two_ranks(built, cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for _, v in got:
        assert v == [3.0, 30.0, 1.0, 1.0]


def _fused_args(built, cuda, n, seed):
    import torch
    from cumicro import fused
    from cumicro.testing import arg_test_distribution, synthetic_states_fused
    CMP = built.CMP if hasattr(built, "CMP") else built
    st = synthetic_states_fused(n, seed=seed)
    d = {k: torch.from_numpy(v).to(cuda) for k, v in st.items()}
    tps = CMP.ThermodynamicsParameters(np.float64)
    blk3 = CMP.pack_icenuc(tps, ad=arg_test_distribution("kappa"), dust=CMP.DustType("Kaolinite"), hom_linear=True)
    return (CMP.Microphysics1MParams(np.float64), CMP.Microphysics2MParams(np.float64), tps, blk3), [d[k] for k in fused.IN_NAMES]


def test_p2p_window_single_rank_is_the_identity_and_the_fused_call_accepts_it(built, cuda):
    """nranks = 1 runs the same exchange code against the rank's own window: sums unchanged, call counter advances, no timeout."""
    import torch
    from cumicro import collective, fused
    win = collective.P2PWindow(0, 1)
    buf = torch.tensor([1.5, -2.0, 3.25, 1e300], dtype=torch.float64, device=cuda)
    for _ in range(5):                                  # both parities, repeatedly
        win.all_reduce(buf)
    torch.cuda.synchronize()
    assert buf.cpu().tolist() == [1.5, -2.0, 3.25, 1e300]
    params, cols = _fused_args(built, cuda, (1 << 16) + 3, seed=4)
    a = fused.fused_1m2m_icenuc(*params, *cols)["diag"].cpu().numpy()
    b = fused.fused_1m2m_icenuc(*params, *cols, p2p_window=win)["diag"].cpu().numpy()
    assert np.array_equal(a, b)
    assert win.status() == (6, 0)
    with pytest.raises(ValueError):
        fused.fused_1m2m_icenuc(*params, *cols, diagnostics=False, p2p_window=win)
    win.destroy()


def _p2p_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import cumicro
    from cumicro import collective, fused
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("gloo", rank=rank, world_size=world)     # only to ship the 64-byte handles
    win = collective.P2PWindow(rank, world).set_timeout(20.0).connect_with_torch_distributed()
    res = []
    for it in range(7):                                                # both parities; a fast rank may run ahead by one call
        buf = torch.tensor([1.0 + rank + it, 10.0 * (rank + 1), 0.5, float(rank) * it], dtype=torch.float64, device=dev)
        win.all_reduce(buf)
        res.append(buf.cpu().numpy().tolist())
    # the fused config-5 call with the exchange in its finish kernel == the slab sums added in rank order
    n_global = (1 << 17) + 5
    lo, hi = fused.slab_bounds(n_global, world, rank)
    params, cols = _fused_args(cumicro.CMP, dev, n_global, seed=11)
    cols = [c[lo:hi].contiguous() for c in cols]
    own = fused.fused_1m2m_icenuc(*params, *cols, reduce=False)["diag"].cpu().numpy()
    dom = fused.fused_1m2m_icenuc(*params, *cols, p2p_window=win)["diag"].cpu().numpy()
    torch.cuda.synchronize()
    box = [None] * world
    dist.all_gather_object(box, own.tolist())
    q.put((rank, res, dom.tolist(), box, win.status()))
    dist.barrier()
    win.destroy()
    dist.destroy_process_group()


def test_p2p_allreduce_and_fused_exchange_two_ranks(built, cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_p2p_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, res, dom, box, status in got:
        for it, v in enumerate(res):
            assert v == [3.0 + 2 * it, 30.0, 1.0, float(it)]
        expect = [box[0][k] + box[1][k] for k in range(4)]            # rank order, one addition: exact
        assert dom == expect
        assert status == (8, 0)
    assert got[0][2] == got[1][2]                                     # bit-identical on both ranks
