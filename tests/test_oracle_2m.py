"""Pins the CPU oracle's 2-moment restatement on the reference's own golden values
(SURVEY.md §8c).  CPU only."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "sb2006_goldens.json")))


def _state(FT):
    s = G["state_gpu"]
    one = lambda v: np.array([v], dtype=FT)
    rho = s["rho"]
    return dict(rho=one(rho), T=one(s["T"]), q_tot=one(s["q_tot"]), q_lcl=one(s["q_lcl"]),
                n_lcl=one(s["N_lcl"] / rho), q_rai=one(s["q_rai"]), n_rai=one(s["N_rai"] / rho))


@pytest.mark.parametrize("limited", [True, False])
def test_sb2006_goldens_f64(built, orc, limited):
    CMP, abi = built.CMP, built._abi
    mp = CMP.Microphysics2MParams(np.float64, is_limited=limited)
    tps = CMP.ThermodynamicsParameters(np.float64)
    st = _state(np.float64)
    out = orc.bmt2m_warm(CMP.pack_2m_warm(mp, tps), *[st[k] for k in ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai")],
                         leaves=True)
    leaf = dict(zip(abi.SB2006_LEAVES, [a[0] for a in out["leaves"]]))
    sb = mp.warm_rain.seifert_beheng
    vt0, vt1 = orc.termvel_2m_rain_sb(sb.pdf_r, CMP.SB2006VelType(np.float64), st["q_rai"], st["rho"],
                                      st["n_rai"] * st["rho"])
    leaf["vt0"], leaf["vt1"] = vt0[0], vt1[0]
    table = dict(G["common"])
    table.update(G["limited" if limited else "notlimited"])
    for name, (val, rtol, where) in table.items():
        got = leaf[name]
        if val == 0.0:
            assert got == 0.0, (name, where)
        else:
            assert abs(got - val) <= rtol * max(abs(val), abs(got)), (name, got, val, where)
    assert leaf["accr_dq_rai"] == -leaf["accr_dq_lcl"]
    # the evaporation literals carry 16 digits: the restatement reproduces them to 1e-13
    ev = G["limited" if limited else "notlimited"]
    assert abs(leaf["evap_dN_rai"] / ev["evap_dN_rai"][0] - 1) < 1e-13
    assert abs(leaf["evap_dq_rai"] / ev["evap_dq_rai"][0] - 1) < 1e-13
    assert abs(leaf["acnv_dq_rai"] / G["common"]["acnv_dq_rai"][0] - 1) < 1e-13


@pytest.mark.parametrize("limited", [True, False])
def test_chen_rain_velocity_golden(built, orc, limited):
    CMP = built.CMP
    g = G["chen_rain_2m"]
    # this unit test of the reference loads toml/SB2006_limiters.toml (microphysics2M_tests.jl:26-31)
    pdf_r = CMP.RainParticlePDF_SB2006(np.float64, is_limited=limited, overrides=CMP.SB2006_LIMITERS_OVERRIDE)
    one = lambda v: np.array([v], dtype=np.float64)
    vt0, vt1 = orc.termvel_2m_rain_chen(pdf_r, CMP.Chen2022VelTypeRain(np.float64), one(g["state"]["q_rai"]),
                                        one(g["state"]["rho"]), one(g["state"]["N_rai"]))
    assert abs(vt0[0] / g["vt0"][0] - 1) < g["vt0"][1]
    assert abs(vt1[0] / g["vt1"][0] - 1) < g["vt1"][1]


def test_cloud_terminal_velocity_closed_form(built, orc):
    """test/microphysics2M_tests.jl:385-416: re-derived closed form and zero gates."""
    import math
    CMP = built.CMP
    pdf_c = CMP.CloudParticlePDF_SB2006(np.float64)
    vel = CMP.StokesRegimeVelType(np.float64)
    rho, q, N = 1.1, 1e-3, 1e8
    one = lambda v: np.array([v], dtype=np.float64)
    vt0, vt1 = orc.termvel_2m_cloud(pdf_c, vel, one(q), one(rho), one(N))
    nu, mu = pdf_c.nu_c, pdf_c.mu_c
    x = rho * q / N
    B = (x * math.gamma((nu + 1) / mu) / math.gamma((nu + 2) / mu)) ** (-mu)
    pref = 2 / 9 * (3 / 4 / math.pi / vel.rho_w) ** (2 / 3) * (vel.rho_w / rho - 1) * vel.grav / vel.nu_air
    M = lambda n: N * B ** (-n / mu) * math.gamma((nu + 1 + n) / mu) / math.gamma((nu + 1) / mu)
    assert abs(vt0[0] / (pref * M(2 / 3) / N) - 1) < 1e-12
    assert abs(vt1[0] / (pref * M(5 / 3) / rho / q) - 1) < 1e-12
    for qq, NN in ((q, 0.0), (0.0, N), (0.0, 0.0)):
        a, b = orc.termvel_2m_cloud(pdf_c, vel, one(qq), one(rho), one(NN))
        assert a[0] == 0 and b[0] == 0


def test_fused_equals_sum_of_leaves(built, orc):
    """BMT:736-779 aggregation order; the reference asserts fused == sum of leaves
    (test/bulk_tendencies_tests.jl:570-610)."""
    CMP, abi = built.CMP, built._abi
    st = built.testing.synthetic_states_2m(4096, seed=3)
    mp = CMP.Microphysics2MParams(np.float64)
    tps = CMP.ThermodynamicsParameters(np.float64)
    out = orc.bmt2m_warm(CMP.pack_2m_warm(mp, tps), *[st[k] for k in ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai")],
                         leaves=True)
    L = dict(zip(abi.SB2006_LEAVES, out["leaves"]))
    rho = st["rho"]
    np.testing.assert_array_equal(out["dq_lcl_dt"], L["cond_dq_lcl"] + L["acnv_dq_lcl"] + L["accr_dq_lcl"])
    np.testing.assert_array_equal(out["dq_rai_dt"], L["evap_dq_rai"] + L["acnv_dq_rai"] + L["accr_dq_rai"])
    dn_r = L["evap_dN_rai"] / rho + L["acnv_dN_rai"] / rho + L["rai_selfcol"] / rho + L["rai_breakup"] / rho + L["numadj_rai"]
    np.testing.assert_allclose(out["dn_rai_dt"], dn_r, rtol=1e-13, atol=0)
    assert np.all(np.isfinite(out["dq_lcl_dt"])) and np.all(np.isfinite(out["dn_rai_dt"]))
    # regimes are exercised by the synthetic states
    assert (L["rai_breakup"] != 0).any() and (L["evap_dq_rai"] < 0).any() and (L["cond_dq_lcl"] > 0).any()


def test_f32_restatement_tracks_f64(built, orc):
    CMP = built.CMP
    st64 = built.testing.synthetic_states_2m(2048, seed=5)
    st32 = {k: v.astype(np.float32) for k, v in st64.items()}
    keys = ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai")
    o64 = orc.bmt2m_warm(CMP.pack_2m_warm(CMP.Microphysics2MParams(np.float64), CMP.ThermodynamicsParameters(np.float64)),
                         *[st32[k].astype(np.float64) for k in keys])
    o32 = orc.bmt2m_warm(CMP.pack_2m_warm(CMP.Microphysics2MParams(np.float32), CMP.ThermodynamicsParameters(np.float32)),
                         *[st32[k] for k in keys])
    for k in ("dq_rai_dt", "dn_lcl_dt"):
        a, b = o32[k].astype(np.float64), o64[k]
        scale = np.maximum(np.abs(b), np.abs(b).mean())
        assert np.median(np.abs(a - b) / scale) < 1e-5


def _alt_cols(name, s, FT):
    one = lambda v: np.array([v], dtype=FT)
    if name.startswith("acnv"):
        return dict(q_lcl=one(s["q_lcl"]), rho=one(s["rho"]), N_d=one(s["N_d"]))
    return dict(q_lcl=one(s["q_lcl"]), q_rai=one(s["q_rai"]), rho=one(s["rho"]))


def test_alternative_closures_goldens(built, orc):
    """KK2000 / B1994 / TC1980 / LD2004 autoconversion and accretion literals (test/gpu_tests.jl:795-818)."""
    g = G["alt_closures"]
    blk = built.CMP.KK2000(np.float64).block
    for name in orc.ALT_2M:
        val, rtol, where = g[name]
        got = orc.alt_2m(blk, name, **_alt_cols(name, g["state"], np.float64))[0]
        assert abs(got / val - 1) <= rtol, (name, got, val, where)
        if rtol < 1e-7:   # the 16-digit literals are reproduced to rounding
            assert abs(got / val - 1) < 1e-14, (name, got, val)
    blk32 = built.CMP.KK2000(np.float32).block
    for name in orc.ALT_2M:
        val = g[name][0]
        got = orc.alt_2m(blk32, name, **_alt_cols(name, g["state"], np.float32))[0]
        assert got.dtype == np.float32 and abs(got / val - 1) < 2e-5, (name, got, val)


def test_alternative_closures_smooth_transition_limits(built, orc):
    """The smooth thresholds approach the sharp ones far from the threshold and stay between 0 and the sharp rate
    (test/microphysics2M_tests.jl:118-192 checks the same property)."""
    blk = built.CMP.B1994(np.float64).block
    rho, q = np.full(4, 1.0), np.full(4, 1e-3)
    for name, N_far in (("acnv_B1994", np.array([1e6, 1e7, 5e9, 1e10])), ("acnv_TC1980", np.array([1e6, 1e7, 1e12, 1e13]))):
        sharp = orc.alt_2m(blk, name, q_lcl=q, rho=rho, N_d=N_far)
        smooth = orc.alt_2m(blk, name, q_lcl=q, rho=rho, N_d=N_far, smooth_transition=True)
        assert np.allclose(smooth, sharp, rtol=1e-3, atol=1e-300), (name, smooth, sharp)
    q = np.array([0.0, 1e-18, 1e-7, 1e-3])
    out = orc.alt_2m(blk, "acnv_LD2004", q_lcl=q, rho=rho, N_d=np.full(4, 1e8))
    assert out[0] == 0 and out[1] == 0 and out[3] > 0
