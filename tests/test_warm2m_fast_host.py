"""The headline kernel body (csrc/cm_sb2006_fast.cuh) is __host__ __device__: its HOST instantiation (same arithmetic; the MUFU
seeds emulated) is checked here against the oracle without a GPU, on the bench inputs and on adversarial ones.  The GPU tests
(tests/test_gpu_2m.py) check the device instantiation on the same criterion."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "warm2m_host.cu")
SO = os.path.join(ROOT, "tests", "native", "_warm2m_host.so")
KEYS = ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai")
OUTS = ("dq_lcl_dt", "dn_lcl_dt", "dq_rai_dt", "dn_rai_dt")


@pytest.fixture(scope="module")
def host(built):
    csrc = os.path.join(ROOT, "cloudmicrophysics.jl_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in ("cm_sb2006_fast.cuh", "cm_sb2006.cuh", "cm_math.cuh", "cm_thermo.cuh", "cm_math_tables.inc")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-fmad=false",
                        "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO, SRC], check=True, capture_output=True)
    return C.CDLL(SO)


def _run(lib, block, st, f32_method=0):
    n = st["rho"].size
    outs = [np.empty(n) for _ in range(4)]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    cols = [np.ascontiguousarray(st[k], dtype=np.float64) for k in KEYS]
    lib.cmt_warm2m_fast(C.byref(block), C.c_int(f32_method), C.c_long(n), *[p(c) for c in cols], *[p(o) for o in outs])
    return dict(zip(OUTS, outs))


def _check(built, orc, lib, block, st, min_forward=0.99):
    from cumicro.testing import assert_parity
    assert lib.cmt_warm2m_supported(C.byref(block)) == 1
    got = _run(lib, block, st)
    cols = [st[k] for k in KEYS]
    ref = orc.bmt2m_warm(block, *cols)
    bound = orc.bmt2m_warm_bound(block, *cols)
    for k in OUTS:
        rep = assert_parity(k, got[k], ref[k], bound=bound[k])
        assert rep["max_rel"] <= 1e-12 and rep["frac_forward_ok"] >= min_forward, (k, rep)


@pytest.mark.parametrize("limited", [True, False])
@pytest.mark.parametrize("number", ["loguniform", "const"])
def test_fast_body_matches_the_oracle_on_the_bench_inputs(built, orc, host, limited, number):
    from cumicro.testing import synthetic_states_2m
    CMP = built.CMP
    st = synthetic_states_2m(1 << 16, seed=1234, number=number)
    mp = CMP.Microphysics2MParams(np.float64, is_limited=limited)
    _check(built, orc, host, CMP.pack_2m_warm(mp, CMP.ThermodynamicsParameters(np.float64)), st)


@pytest.mark.parametrize("limited", [True, False])
def test_fast_body_adversarial_inputs(built, orc, host, limited):
    """Tiny rain with leftover number (xr_mean / xr_min down to 1e-16: the not-limited t* leaves exp's fast domain),
    huge mean drops (Dr far above Deq), negative / zero inputs (the BMT:827-836 clamps), values straddling eps."""
    from cumicro.testing import synthetic_states_2m
    CMP = built.CMP
    n = 1 << 15
    st = synthetic_states_2m(n, seed=77)
    rng = np.random.Generator(np.random.PCG64(5))
    xr_min = 2.6e-10
    ratio = 10.0 ** rng.uniform(-16, 6, n)                     # xr_mean / xr_min
    st["q_rai"] = 10.0 ** rng.uniform(-18, -3, n)
    st["n_rai"] = st["q_rai"] / (ratio * xr_min)
    st["q_lcl"] = 10.0 ** rng.uniform(-18, -2.5, n)
    st["n_lcl"] = 10.0 ** rng.uniform(0, 10, n)
    eps = np.finfo(np.float64).eps
    for k in ("q_rai", "q_lcl", "n_rai", "n_lcl"):
        idx = rng.choice(n, n // 16, replace=False)
        st[k][idx[: n // 64]] = 0.0
        st[k][idx[n // 64: n // 32]] = -st[k][idx[n // 64: n // 32]]
        st[k][idx[n // 32: 3 * n // 64]] = eps * rng.uniform(0.5, 2.0, n // 64)
        st[k][idx[3 * n // 64:]] = eps
    st["q_tot"] = np.maximum(st["q_tot"], 0) + np.abs(st["q_lcl"]) + np.abs(st["q_rai"])
    st["q_tot"][rng.choice(n, 64, replace=False)] *= -1.0
    mp = CMP.Microphysics2MParams(np.float64, is_limited=limited)
    _check(built, orc, host, CMP.pack_2m_warm(mp, CMP.ThermodynamicsParameters(np.float64)), st, min_forward=0.97)


def test_fast_body_with_non_default_parameter_values(built, orc, host):
    """The fast body specialises on the STRUCTURE of the default block (exponents 3 / 4 / -5); every value stays a run-time
    parameter, including evap.rho0 != accr.rho0 != pdf_r.rho0 (folded into host constants)."""
    from cumicro.testing import synthetic_states_2m
    CMP = built.CMP
    st = synthetic_states_2m(1 << 15, seed=3)
    mp = CMP.Microphysics2MParams(np.float64)
    blk = CMP.pack_2m_warm(mp, CMP.ThermodynamicsParameters(np.float64))
    blk.sb.evap.rho0 = 1.1; blk.sb.accr.rho0 = 1.3; blk.sb.pdf_r.rho0 = 1.225
    blk.sb.acnv.a = 0.71; blk.sb.acnv.A = 350.0; blk.sb.accr.tau0 = 4e-4; blk.sb.self.kappa_rr = 55.0
    blk.sb.evap.beta_vent_0 = -0.4; blk.sb.evap.alpha = 150.0; blk.sb.evap.beta = 0.62
    blk.sb.brek.kappa_br = 1800.0; blk.sb.pdf_r.lam_max = 8000.0; blk.sb.pdf_r.N0_min = 5e5
    blk.condevap_tau_relax = 7.0; blk.sb.numadj_tau = 50.0
    _check(built, orc, host, blk, st)
    # a block WITHOUT the default structure is refused by the fast path (the library then runs the general body)
    blk.sb.acnv.b = 2.5
    assert host.cmt_warm2m_supported(C.byref(blk)) == 0


def test_log_abs_accuracy(host):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    rng = np.random.default_rng(1)
    x = np.concatenate([10 ** rng.uniform(-300, 300, 1500), rng.uniform(0.5, 2, 1500), 1 + rng.uniform(-1e-2, 1e-2, 300)])
    y = np.empty_like(x)
    host.cmt_log_abs(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_long(x.size))
    worst = 0.0
    for xi, yi in zip(x, y):
        t = mp.log(mp.mpf(float(xi)))
        worst = max(worst, float(abs(mp.mpf(float(yi)) - t) / max(1, abs(t))))
    assert worst < 2.5e-16, worst   # ABSOLUTE accuracy ~1.5e-16 max(1, |log x|)
