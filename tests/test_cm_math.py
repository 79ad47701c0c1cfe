"""Accuracy of the hand-written Float64 device math (cm_math.cuh), measured on the HOST
instantiation of the same code against mpmath (40 digits).  The GPU instantiation differs
only in the hardware seeds (MUFU.RCP64H / lg2 / ex2), which the Newton steps erase."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

mp = pytest.importorskip("mpmath")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "cm_math_host.cu")
SO = os.path.join(ROOT, "tests", "native", "_cm_math_host.so")


@pytest.fixture(scope="module")
def lib():
    deps = [SRC, os.path.join(ROOT, "cloudmicrophysics.jl_b200", "csrc", "cm_math.cuh"),
            os.path.join(ROOT, "cloudmicrophysics.jl_b200", "csrc", "cm_math_tables.inc")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-gencode",
                        "arch=compute_100a,code=sm_100a", "-o", SO, SRC], check=True, capture_output=True)
    return C.CDLL(SO)


def _run(lib, fn, x):
    y = np.empty_like(x)
    getattr(lib, fn)(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_long(x.size))
    return y


def _max_ulp(y, x, f):
    mp.mp.dps = 40
    worst = 0.0
    for xi, yi in zip(x, y):
        t = f(mp.mpf(float(xi)))
        worst = max(worst, float(abs((mp.mpf(float(yi)) - t) / t) / mp.mpf(2) ** -52))
    return worst


def test_exp(lib):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-700, 700, 1500), rng.uniform(-2, 2, 1500), rng.uniform(-1e-3, 1e-3, 300), [0.0]])
    assert _max_ulp(_run(lib, "cmt_exp", x), x, mp.exp) < 1.6   # relative error in units of 2^-52
    y = _run(lib, "cmt_exp_full", np.array([-800.0, 800.0, np.nan, 0.0, -745.0, 709.5, -720.0, -744.0, 709.9]))
    assert y[0] == 0 and np.isinf(y[1]) and np.isnan(y[2]) and y[3] == 1 and np.isinf(y[8])
    # gradual underflow into the subnormals and the top of the range, like libm
    for got, x in zip(y[[4, 5, 6, 7]], (-745.0, 709.5, -720.0, -744.0)):
        assert abs(got - np.exp(x)) <= max(2e-15 * np.exp(x), 5e-324), (x, got, np.exp(x))


def test_log(lib):
    rng = np.random.default_rng(1)
    x = np.concatenate([10 ** rng.uniform(-300, 300, 1500), rng.uniform(0.5, 2, 1500), 1 + rng.uniform(-1e-2, 1e-2, 600),
                        1 + rng.uniform(-1e-6, 1e-6, 300)])
    assert _max_ulp(_run(lib, "cmt_log", x), x, mp.log) < 2.0
    assert _run(lib, "cmt_log", np.array([1.0]))[0] == 0.0


def test_cbrt_and_rcp(lib):
    rng = np.random.default_rng(2)
    x = np.concatenate([10 ** rng.uniform(-300, 300, 1500), rng.uniform(0.5, 16, 1500)])
    assert _max_ulp(_run(lib, "cmt_cbrt", x), x, mp.cbrt) < 1.0
    assert _max_ulp(_run(lib, "cmt_rcp", x), x, lambda t: 1 / t) < 1.0
    assert _max_ulp(_run(lib, "cmt_rcbrt", x), x, lambda t: 1 / mp.cbrt(t)) < 2.5   # the by-product x^(-1/3) of cbrt_pair_
    assert _max_ulp(_run(lib, "cmt_sqrt", x), x, mp.sqrt) < 1.0


def test_pow(lib):
    """powp_(x, y) = exp_(y logp_(x)): relative error <= ~2 ulp * (|y ln x| + 1)."""
    mp.mp.dps = 40
    rng = np.random.default_rng(3)
    xs = 10 ** rng.uniform(-12, 3, 1500)
    ps = rng.uniform(-5, 5, 1500)
    y = np.empty_like(xs)
    lib.cmt_pow(xs.ctypes.data_as(C.c_void_p), ps.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_long(xs.size))
    worst = 0.0
    for a, b, c in zip(xs, ps, y):
        t = mp.power(mp.mpf(float(a)), mp.mpf(float(b)))
        amp = abs(mp.mpf(float(b)) * mp.log(mp.mpf(float(a)))) + 1
        worst = max(worst, float(abs((mp.mpf(float(c)) - t) / t) / amp / mp.mpf(2) ** -52))
    assert worst < 2.0


def test_division_through_the_shared_reciprocal_is_ieee(lib):
    """divr_(x, d, rcp_cr_(d)) equals the IEEE quotient x / d bit for bit (Markstein), including x = 0 — the case CUDA's
    own division sends to its slow path, and the reason the linearised 1-moment step uses this form."""
    rng = np.random.default_rng(5)
    n = 200000
    d = np.concatenate([10 ** rng.uniform(-12, 8, n // 2), rng.uniform(1e-10, 1.0, n // 2)])
    x = np.concatenate([rng.normal(size=n // 2) * 10 ** rng.uniform(-30, 5, n // 2), rng.uniform(-1e-3, 1e-3, n // 2)])
    x[::50] = 0.0
    y = np.empty_like(x)
    lib.cmt_divr(x.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_long(n))
    ref = x / d
    assert np.all(y[x == 0] == 0)
    assert np.mean(y == ref) > 0.9999, np.mean(y == ref)
    assert np.max(np.abs(y - ref) / np.maximum(np.abs(ref), 1e-300)) < 2.3e-16


def test_log1p_of_a_positive_argument(lib):
    """log1p_pos_ (Kahan's log(u) y/(u-1) on the table-driven log) over the range log1pexp feeds it: e^x, x in [-36.7, 18.02]."""
    rng = np.random.default_rng(6)
    y = np.concatenate([np.exp(rng.uniform(-36.7, 18.02, 4000)), 10 ** rng.uniform(-17, -14, 300), [1e-300, 1.0, 2.0 ** -53, 2.0 ** -52]])
    assert _max_ulp(_run(lib, "cmt_log1p_pos", y), y, mp.log1p) < 4.5



def test_erf_fast_absolute_error(lib):
    """erf_fast_ = 1 - exp_(-|x| Q): ONE branch-free piece, |error| < 2.5 units of 2^-53 ABSOLUTE (its use, N (1 - erf u) / 2, needs
    no more: AA:256), odd, saturating at +-1 beyond 6, NaN kept."""
    mp.mp.dps = 40
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-6.5, 6.5, 4000), rng.uniform(-1, 1, 2000), 10.0 ** rng.uniform(-300, 0, 500), np.linspace(0, 6, 1201)])
    y = _run(lib, "cmt_erf_fast", x)
    worst = max(float(abs(mp.mpf(float(b)) - mp.erf(mp.mpf(float(a)))) / mp.mpf(2) ** -53) for a, b in zip(x, y))
    assert worst < 2.5, worst
    assert np.array_equal(_run(lib, "cmt_erf_fast", -x), -y)
    s = _run(lib, "cmt_erf_fast", np.array([0.0, 7.0, -7.0, np.inf, -np.inf, np.nan, 1e300, -0.0]))
    assert s[0] == 0 and s[1] == 1 and s[2] == -1 and s[3] == 1 and s[4] == -1 and np.isnan(s[5]) and s[6] == 1 and s[7] == 0
