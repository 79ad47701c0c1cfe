"""GPU parity of the cloud diagnostics (src/CloudDiagnostics.jl) through the C-ABI vs the CPU oracle."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "diag_goldens.json")))


def _dev(cols, dev, dtype=None):
    import torch
    return [torch.from_numpy(np.ascontiguousarray(c if dtype is None else c.astype(dtype))).to(dev) for c in cols]


def test_reference_values_through_the_gpu(built, cuda):
    CMP, CMD = built.CMP, built.CMD
    g = G["sb2006_2m"]
    cols = [np.array(g[k]) for k in ("q_lcl", "q_rai", "N_lcl", "N_rai")] + [np.ones(5)]
    for FT, za, ra in ((np.float64, g["Z_atol"], g["reff_atol"]), (np.float32, 1e-3, 1e-6)):
        for limited in (True, False):
            sb = CMP.SB2006(FT, is_limited=limited, overrides=CMP.SB2006_LIMITERS_OVERRIDE)
            Z, reff = CMD.radar_reflectivity_and_effective_radius_2M(sb, *_dev(cols, cuda, FT))
            assert np.all(np.abs(Z.cpu().numpy() - np.array(g["Z"])) <= za), (FT, limited, Z)
            assert np.all(np.abs(reff.cpu().numpy() - np.array(g["reff"])) <= ra), (FT, limited, reff)
            assert np.array_equal(CMD.radar_reflectivity_2M(sb, *_dev(cols, cuda, FT)).cpu().numpy(), Z.cpu().numpy())
            assert np.array_equal(CMD.effective_radius_2M(sb, *_dev(cols, cuda, FT)).cpu().numpy(), reff.cpu().numpy())
    mp, tps = CMP.Microphysics1MParams(np.float64), CMP.ThermodynamicsParameters(np.float64)
    g1 = G["radar_1m"]
    q = np.array([c[0] for c in g1["cases"]])
    Z = CMD.radar_reflectivity_1M(mp, tps, *_dev([q, np.full(2, g1["rho"])], cuda)).cpu().numpy()
    for z, (_, val, atol, where) in zip(Z, g1["cases"]):
        assert abs(z - val) <= atol, where
    gl = G["liu_hallett"]
    one = lambda v: np.array([v])
    r = CMD.effective_radius_Liu_Hallet_97(mp.block.cloud_liquid, *_dev([one(gl["rho"]), one(gl["q_lcl"]), one(gl["N_lcl"]), one(gl["q_rai"]), one(gl["N_rai"])], cuda))
    assert abs(float(r[0]) - gl["reff"]) <= gl["atol"]
    assert CMD.effective_radius_const(mp.block.cloud_liquid) == G["const"]["cloud_liquid"]
    assert CMD.effective_radius_const(mp.block.cloud_ice) == G["const"]["cloud_ice"]


@pytest.mark.parametrize("limited", [True, False])
def test_diag_2m_parity(built, orc, cuda, limited):
    """Seeded 2-moment states: Z within 1e-11 dBZ absolute (it crosses zero) + 1e-12 relative, r_eff within 1e-12 relative;
    clipped (-150 dBZ) and zero (r_eff) patterns identical.  Float32 method: 4 ULP (+ 2e-6 dBZ for Z near its zero crossing)."""
    from cumicro.testing import synthetic_states_2m, assert_f32_method
    CMP, CMD = built.CMP, built.CMD
    n = 1 << 17
    st = synthetic_states_2m(n, seed=77)
    cols = [st["q_lcl"], st["q_rai"], st["n_lcl"] * st["rho"], st["n_rai"] * st["rho"], st["rho"]]
    sb = CMP.SB2006(np.float64, is_limited=limited)
    Zg, rg = [t.cpu().numpy() for t in CMD.radar_reflectivity_and_effective_radius_2M(sb, *_dev(cols, cuda))]
    Zr, rr = orc.diag_2m(sb.pdf_c, sb.pdf_r, *cols)
    assert np.array_equal(Zg == -150.0, Zr == -150.0) and np.array_equal(rg == 0, rr == 0)
    assert np.all(np.abs(Zg - Zr) <= 1e-11 + 1e-12 * np.abs(Zr)), np.abs(Zg - Zr).max()
    nz = rr != 0
    assert np.abs(rg[nz] / rr[nz] - 1).max() <= 1e-12
    assert (Zr > -150).mean() > 0.5 and nz.mean() > 0.5
    # Float32 method
    c32 = [c.astype(np.float32) for c in cols]
    sb32 = CMP.SB2006(np.float32, is_limited=limited)
    Z32, r32 = [t.cpu().numpy() for t in CMD.radar_reflectivity_and_effective_radius_2M(sb32, *_dev(c32, cuda))]
    with orc.f32_thresholds():
        Zt, rt = orc.diag_2m(CMP.widen(sb32.pdf_c), CMP.widen(sb32.pdf_r), *[c.astype(np.float64) for c in c32])
    assert np.array_equal(Z32 == -150.0, Zt == -150.0)
    assert np.all(np.abs(Z32.astype(np.float64) - Zt) <= 2e-6 + 4 * np.spacing(np.abs(Zt).astype(np.float32)))
    assert_f32_method("r_eff", r32, rt.astype(np.float32), rt)


def test_diag_1m_and_liu_hallett_parity(built, orc, cuda):
    CMP, CMD = built.CMP, built.CMD
    rng = np.random.Generator(np.random.PCG64(5))
    n = 1 << 16
    q = 10.0 ** rng.uniform(-9, -2, n)
    q[::50] = 0.0
    q[1::50] = -1e-6
    rho = rng.uniform(0.2, 1.3, n)
    mp, tps = CMP.Microphysics1MParams(np.float64), CMP.ThermodynamicsParameters(np.float64)
    Zg = CMD.radar_reflectivity_1M(mp, tps, *_dev([q, rho], cuda)).cpu().numpy()
    Zr = orc.diag_1m(CMP.pack_1m(mp, tps), q, rho)
    assert np.all(np.abs(Zg - Zr) <= 1e-11 + 1e-12 * np.abs(Zr)), np.abs(Zg - Zr).max()
    N_l, N_r, q_r = 10.0 ** rng.uniform(6, 9, n), 10.0 ** rng.uniform(2, 6, n), 10.0 ** rng.uniform(-8, -3, n)
    N_l[::64] = 0.0
    N_r[::64] = 0.0
    qa = np.abs(q)
    got = CMD.effective_radius_Liu_Hallet_97(1000.0, *_dev([rho, qa, N_l, q_r, N_r], cuda)).cpu().numpy()
    ref = orc.diag_reff_lh97(1000.0, rho, qa, N_l, q_r, N_r)
    assert np.array_equal(got == 0, ref == 0) and np.abs(got[ref != 0] / ref[ref != 0] - 1).max() <= 1e-12
    got3 = CMD.effective_radius_Liu_Hallet_97(mp.block.cloud_liquid, *_dev([rho, qa], cuda)).cpu().numpy()
    ref3 = orc.diag_reff_lh97(1000.0, rho, qa)
    assert np.abs(got3[ref3 != 0] / ref3[ref3 != 0] - 1).max() <= 1e-12
    with pytest.raises(TypeError):
        CMD.effective_radius_Liu_Hallet_97(1000.0, *_dev([rho, qa, N_l], cuda))


@pytest.mark.parametrize("n", [0, 1, 33, 1000])
def test_leaf_entry_points_ragged_sizes_and_misaligned_columns(built, orc, cuda, n):
    """Empty, tiny and ragged columns, and columns that start 8 bytes off a 16-byte boundary (the scalar-access path),
    through the diagnostics and the alternative-closure entry points: same bits as the aligned call."""
    import torch
    CMP, CMD, CM2 = built.CMP, built.CMD, built.CM2
    rng = np.random.Generator(np.random.PCG64(n + 1))
    mk = lambda lo, hi: torch.from_numpy(10.0 ** rng.uniform(lo, hi, n + 1)).to(cuda)
    q_l, q_r, N_l, N_r, rho = mk(-6, -3), mk(-7, -3), mk(6, 9), mk(2, 6), mk(-0.5, 0.1)
    sb = CMP.SB2006(np.float64)
    al = [t[:n].clone() for t in (q_l, q_r, N_l, N_r, rho)]          # aligned copies
    mis = [t[1:n + 1] for t in (q_l, q_r, N_l, N_r, rho)]            # views at +8 bytes
    al_of_mis = [t.clone() for t in mis]
    Z, r = CMD.radar_reflectivity_and_effective_radius_2M(sb, *al)
    assert Z.shape == (n,) and r.shape == (n,)
    Zm, rm = CMD.radar_reflectivity_and_effective_radius_2M(sb, *mis)
    Za, ra = CMD.radar_reflectivity_and_effective_radius_2M(sb, *al_of_mis)
    assert torch.equal(Zm, Za) and torch.equal(rm, ra)
    if n:
        Zr, rr = orc.diag_2m(sb.pdf_c, sb.pdf_r, *[t.cpu().numpy() for t in al])
        assert np.all(np.abs(Z.cpu().numpy() - Zr) <= 1e-11 + 1e-12 * np.abs(Zr))
    kk = CMP.KK2000(np.float64)
    a = CM2.conv_q_lcl_to_q_rai(kk, mis[0], mis[4], mis[2])
    b = CM2.conv_q_lcl_to_q_rai(kk, al_of_mis[0], al_of_mis[4], al_of_mis[2])
    assert a.shape == (n,) and torch.equal(a, b)
    lh = CMD.effective_radius_Liu_Hallet_97(1000.0, mis[4], mis[0])
    assert lh.shape == (n,) and torch.equal(lh, CMD.effective_radius_Liu_Hallet_97(1000.0, al_of_mis[4], al_of_mis[0]))
