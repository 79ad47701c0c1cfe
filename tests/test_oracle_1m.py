"""Pins the CPU oracle's 1-moment / non-equilibrium restatement on the reference's own
golden values and checks the fused-path identities the reference asserts
algebraically (test/bulk_tendencies_tests.jl).  CPU only."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "m1_goldens.json")))
one = lambda v, FT=np.float64: np.array([v], dtype=FT)


def _blk(built, FT=np.float64, **opts):
    CMP = built.CMP
    mp = CMP.Microphysics1MParams(FT, **opts)
    return mp, CMP.pack_1m(mp, CMP.ThermodynamicsParameters(FT))


def test_accretion_goldens(built, orc):
    mp, blk = _blk(built)
    s = G["accretion"]["state"]
    q = one(s["q"])
    cold = orc.bmt1m(blk, one(s["rho"]), one(263.0), one(15e-3), q, q, q, q, mode="verbose")
    warm = orc.bmt1m(blk, one(s["rho"]), one(283.0), one(15e-3), q, q, q, q, mode="verbose")
    for k, (val, where) in G["accretion"].items():
        if k == "state":
            continue
        got = (warm if k.endswith("_warm") else cold)[k][0]
        assert abs(got / val - 1) < 2e-14, (k, got, val, where)


def test_snow_melt_and_velocity_goldens(built, orc):
    CMP = built.CMP
    mp, blk = _blk(built)
    g = G["snow_melt"]
    o = orc.bmt1m(blk, one(g["rho"]), one(273.15 + g["dT"]), one(1e-2), one(0.0), one(0.0), one(0.0), one(g["q_sno"]), mode="verbose")
    assert abs(o["S_melt_sno_rai"][0] / g["value"] - 1) < 2e-14
    vels = {"rain_chen": CMP.Chen2022VelTypeRain, "snow_chen": CMP.Chen2022VelTypeLargeIce,
            "cloud_liquid_stokes": CMP.StokesRegimeVelType, "cloud_ice_chen": CMP.Chen2022VelTypeSmallIce}
    for kind, rho, q, val, where in G["velocities"]:
        got = orc.termvel_1m(blk, kind, one(rho), one(q), vels[kind](np.float64))[0]
        assert abs(got / val - 1) < 1e-14, (kind, got, val, where)


def test_noneq_goldens(built, orc):
    from cumicro.testing import psat_liq, psat_ice
    mp, blk = _blk(built)
    g = G["noneq"]
    rho, T = g["rho"], g["T"]
    Rv = built.CMP.DEFAULTS["gas_constant_vapor"]
    for key, psat, want in (("S_phase_change_vap_lcl", psat_liq, g["cond"]), ("S_phase_change_vap_icl", psat_ice, g["dep"])):
        q_tot = 1.2 * psat(T) / (Rv * rho * T)
        o = orc.bmt1m(blk, one(rho), one(T), one(q_tot), one(0.0), one(0.0), one(0.0), one(0.0), mode="verbose")
        assert abs(o[key][0] / want - 1) < 1e-6, (key, o[key][0], want)   # the reference's own rtol
        assert abs(o[key][0] / want - 1) < 1e-13                           # and in fact to all printed digits


def test_option_variants_goldens(built, orc):
    from cumicro.testing import psat_ice
    CMP = built.CMP
    mp, blk = _blk(built, rain_autoconversion=CMP.PrescribedNd())
    g = G["prescribed_nd"]
    o = orc.bmt1m(blk, one(1.0), one(280.0), one(0.0), one(g["q_lcl"]), one(0.0), one(0.0), one(0.0), mode="verbose")
    assert abs(o["S_acnv_lcl_rai"][0] / g["value"] - 1) < g["rtol"]
    mp, blk = _blk(built, snow_autoconversion=CMP.WithSupersaturation())
    rho, T = 1.0, 273.15 - 10
    qv = 1.02 * psat_ice(T) / (CMP.DEFAULTS["gas_constant_vapor"] * rho * T)
    qi = 0.03 * qv
    o = orc.bmt1m(blk, one(rho), one(T), one(qv + qi + 2e-4), one(0.0), one(qi), one(1e-4), one(1e-4), mode="verbose")
    assert abs(o["S_acnv_icl_sno"][0] / G["snow_acnv_with_supersat"]["value"] - 1) < 1.5e-8


def test_disabled_processes_and_aggregation(built, orc):
    """`nothing` options zero their terms; Instantaneous == aggregation of Verbose terms
    (BMT:227-252) bit for bit; total water only moves through the vapour terms."""
    CMP = built.CMP
    st = built.testing.synthetic_states_1m(4096, seed=3)
    cols = [st[k] for k in ("rho", "T", "q_tot", "q_lcl", "q_icl", "q_rai", "q_sno")]
    mp, blk = _blk(built)
    v = orc.bmt1m(blk, *cols, mode="verbose")
    i = orc.bmt1m(blk, *cols, mode="instantaneous")
    for k in orc.OUT_1M:
        assert np.array_equal(v[k], i[k])
    tot = v["dq_lcl_dt"] + v["dq_icl_dt"] + v["dq_rai_dt"] + v["dq_sno_dt"]
    vap = v["S_phase_change_vap_lcl"] + v["S_phase_change_vap_icl"] + v["S_phase_change_vap_rai"] + v["S_phase_change_vap_sno"]
    scale = sum(np.abs(v[k]) for k in orc.SRC_1M)
    assert np.all(np.abs(tot - vap) <= 1e-14 * scale + 1e-300)     # condensate is conserved by collisions / melting
    assert (v["S_accr_rai_sno_cold"] > 0).any() and (v["S_accr_rai_sno_warm"] > 0).any() and (v["S_melt_sno_rai"] > 0).any()
    mp, blk = _blk(built, rain_snow_accretion=None, snow_melt=None, cloud_ice_formation=None)
    z = orc.bmt1m(blk, *cols, mode="verbose")
    for k in ("S_accr_rai_sno_cold", "S_accr_rai_sno_warm", "S_accr_melt_rai_sno", "S_melt_sno_rai", "S_phase_change_vap_icl"):
        assert not z[k].any()
    assert np.array_equal(z["S_accr_lcl_rai"], v["S_accr_lcl_rai"])


def test_linearized_average_properties(built, orc):
    """test/bulk_tendencies_tests.jl:886-977: LinearizedAverage -> Instantaneous as dt -> 0;
    implicit steps never drive a species negative."""
    st = built.testing.synthetic_states_1m(2048, seed=9)
    cols = [st[k] for k in ("rho", "T", "q_tot", "q_lcl", "q_icl", "q_rai", "q_sno")]
    mp, blk = _blk(built)
    inst = orc.bmt1m(blk, *cols)
    small = orc.bmt1m(blk, *cols, mode="linearized_average", dt=1e-6, nsub=1)
    for k in orc.OUT_1M:
        scale = np.abs(inst[k]).max()
        # the vapour cap (alpha) and q_min floors make this an approximation, not an identity
        assert np.median(np.abs(small[k] - inst[k])) <= 1e-3 * scale
    for dt, nsub in ((60.0, 1), (600.0, 3)):
        avg = orc.bmt1m(blk, *cols, mode="linearized_average", dt=dt, nsub=nsub)
        for k, q in zip(orc.OUT_1M, ("q_lcl", "q_icl", "q_rai", "q_sno")):
            assert np.all(np.isfinite(avg[k]))
            assert np.all(st[q] + dt * avg[k] >= -1e-12)


def test_0m_remove_precipitation(built, orc):
    """test/microphysics0M_tests.jl:12-55: no removal without cloud, the qc_0 and S_0 thresholds, input clamps (BMT:662-663)."""
    CMP = built.CMP
    for FT in (np.float64, np.float32):
        p = CMP.Parameters0M(FT)
        tau, qc0, S0 = FT(p.tau_precip), FT(p.qc_0), FT(p.S_0)
        z = np.zeros(3, FT)
        assert np.all(orc.bmt0m(p, z, z) == 0) and np.all(orc.bmt0m(p, z, z, np.full(3, 10e-3, FT)) == 0)
        qc = FT(3e-3)
        lf = np.array([0, 0.5, 1.0], FT)
        ql, qi = qc * lf, (FT(1) - lf) * qc
        assert np.array_equal(orc.bmt0m(p, ql, qi), -np.maximum(FT(0), ql + qi - qc0) / tau)
        qvs = np.full(3, 10e-3, FT)
        assert np.array_equal(orc.bmt0m(p, ql, qi, qvs), -np.maximum(FT(0), ql + qi - S0 * qvs) / tau)
        assert np.array_equal(orc.bmt0m(p, -ql, qi), orc.bmt0m(p, z, qi))
