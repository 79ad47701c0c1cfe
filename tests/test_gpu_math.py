"""Accuracy of the device math on the GPU itself (hardware MUFU seeds) vs mpmath."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
mp = pytest.importorskip("mpmath")


def _probe(built, cuda, fn, x, y=None):
    import torch
    xd = torch.from_numpy(x).to(cuda)
    yd = torch.from_numpy(y if y is not None else np.zeros_like(x)).to(cuda)
    out = torch.empty_like(xd)
    lib = built._abi.load()
    st = lib.cumicro_probe_math_f64(fn, C.c_int64(x.size), C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr()),
                                    C.c_void_p(out.data_ptr()), None)
    built._abi.check(st, "cumicro_probe_math_f64")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _max_ulp(y, x, f):
    mp.mp.dps = 40
    worst = 0.0
    for xi, yi in zip(x, y):
        t = f(mp.mpf(float(xi)))
        worst = max(worst, float(abs((mp.mpf(float(yi)) - t) / t) / mp.mpf(2) ** -52))
    return worst


def test_device_math_accuracy(built, cuda):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-700, 700, 2000), rng.uniform(-2, 2, 2000), rng.uniform(-1e-3, 1e-3, 400)])
    assert _max_ulp(_probe(built, cuda, 0, x), x, mp.exp) < 1.6
    x = np.concatenate([10 ** rng.uniform(-300, 300, 2000), rng.uniform(0.5, 2, 2000), 1 + rng.uniform(-1e-2, 1e-2, 600)])
    assert _max_ulp(_probe(built, cuda, 1, x), x, mp.log) < 2.0
    x = np.concatenate([10 ** rng.uniform(-300, 300, 2000), rng.uniform(0.5, 16, 2000)])
    assert _max_ulp(_probe(built, cuda, 2, x), x, mp.cbrt) < 1.0
    assert _max_ulp(_probe(built, cuda, 3, x), x, lambda t: 1 / t) < 1.0
    assert _max_ulp(_probe(built, cuda, 6, x), x, mp.sqrt) < 1.0
    xs, ps = 10 ** rng.uniform(-12, 3, 2000), rng.uniform(-5, 5, 2000)
    got = _probe(built, cuda, 4, xs, ps)
    mp.mp.dps = 40
    worst = 0.0
    for a, b, c in zip(xs, ps, got):
        t = mp.power(mp.mpf(float(a)), mp.mpf(float(b)))
        worst = max(worst, float(abs((mp.mpf(float(c)) - t) / t) / (abs(mp.mpf(float(b)) * mp.log(mp.mpf(float(a)))) + 1) / mp.mpf(2) ** -52))
    assert worst < 2.0
    xs = np.array([-800.0, 800.0, np.nan, 0.0, -745.0, 709.5, -720.0, -744.0, 709.9])
    y = _probe(built, cuda, 5, xs)
    assert y[0] == 0 and np.isinf(y[1]) and np.isnan(y[2]) and y[3] == 1 and np.isinf(y[8])
    for got, x in zip(y[[4, 5, 6, 7]], (-745.0, 709.5, -720.0, -744.0)):
        assert abs(got - np.exp(x)) <= max(2e-15 * np.exp(x), 5e-324), (x, got, np.exp(x))


def test_device_erf_log1p_rcbrt_division(built, cuda):
    """The functions added for the ARG2000 / 1-moment kernels, on the device: erf_ (absolute error in units of 2^-53),
    log1p_pos_, the reciprocal cube root, and the shared-reciprocal division (bit-identical to IEEE)."""
    mp.mp.dps = 40
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.uniform(-6.2, 6.2, 3000), rng.uniform(-0.9, 0.9, 1000)])
    y = _probe(built, cuda, 7, x)
    worst = max(float(abs(mp.mpf(float(b)) - mp.erf(mp.mpf(float(a)))) / mp.mpf(2) ** -53) for a, b in zip(x, y))
    assert worst < 2.0, worst
    s = _probe(built, cuda, 7, np.array([0.0, 7.0, -7.0, np.inf, -np.inf, np.nan]))
    assert s[0] == 0 and s[1] == 1 and s[2] == -1 and s[3] == 1 and s[4] == -1 and np.isnan(s[5])
    # erf_fast_ (the ARG2000 kernel's): absolute error, printed so that a failure shows the value
    yf = _probe(built, cuda, 11, x)
    worst_f = max(float(abs(mp.mpf(float(b)) - mp.erf(mp.mpf(float(a)))) / mp.mpf(2) ** -53) for a, b in zip(x, yf))
    assert worst_f < 3.0, worst_f
    sf = _probe(built, cuda, 11, np.array([0.0, 7.0, -7.0, np.inf, -np.inf, np.nan]))
    assert sf[0] == 0 and sf[1] == 1 and sf[2] == -1 and sf[3] == 1 and sf[4] == -1 and np.isnan(sf[5])
    v = np.exp(rng.uniform(-36.7, 18.02, 3000))
    assert _max_ulp(_probe(built, cuda, 8, v), v, mp.log1p) < 4.5
    w = 10 ** rng.uniform(-300, 300, 2000)
    assert _max_ulp(_probe(built, cuda, 9, w), w, lambda t: 1 / mp.cbrt(t)) < 2.5
    d = 10 ** rng.uniform(-12, 8, 100000)
    a = rng.normal(size=d.size) * 10 ** rng.uniform(-30, 5, d.size)
    a[::50] = 0.0
    q = _probe(built, cuda, 10, a, d)
    assert np.mean(q == a / d) > 0.9999 and np.all(q[a == 0] == 0)
