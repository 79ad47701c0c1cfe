"""Pins the CPU oracle's restatement of src/CloudDiagnostics.jl on the reference's own test values
(test/cloud_diagnostics.jl).  CPU only."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "diag_goldens.json")))


def test_radar_reflectivity_1m(built, orc):
    CMP = built.CMP
    blk = CMP.pack_1m(CMP.Microphysics1MParams(np.float64), CMP.ThermodynamicsParameters(np.float64))
    g = G["radar_1m"]
    for q, val, atol, where in g["cases"]:
        got = orc.diag_1m(blk, np.array([q]), np.array([g["rho"]]))[0]
        assert abs(got - val) <= atol, (got, val, where)
    # clipped at -150 dBZ for no rain; monotone in q
    z = orc.diag_1m(blk, np.array([0.0, 1e-8, 1e-6, 1e-4, 1e-3]), np.ones(5))
    assert np.all(np.diff(z) > 0) and z[0] >= -150


def test_sb2006_reflectivity_and_effective_radius(built, orc):
    CMP = built.CMP
    g = G["sb2006_2m"]
    cols = [np.array(g[k]) for k in ("q_lcl", "q_rai", "N_lcl", "N_rai")] + [np.ones(5)]
    for limited in (True, False):
        sb = CMP.SB2006(np.float64, is_limited=limited, overrides=CMP.SB2006_LIMITERS_OVERRIDE)
        Z, reff = orc.diag_2m(sb.pdf_c, sb.pdf_r, *cols)
        assert np.all(np.abs(Z - np.array(g["Z"])) <= g["Z_atol"]), (limited, Z)
        assert np.all(np.abs(reff - np.array(g["reff"])) <= g["reff_atol"]), (limited, reff)
        assert abs(Z[0] - g["Z"][0]) < 1e-11       # the one 16-digit literal is reproduced to 13 digits (1.6e-12 dBZ)
        assert np.all(Z[2:] == -150.0) and np.all(reff[3:] == 0.0)
        sb32 = CMP.SB2006(np.float32, is_limited=limited, overrides=CMP.SB2006_LIMITERS_OVERRIDE)
        Z32, r32 = orc.diag_2m(sb32.pdf_c, sb32.pdf_r, *cols)
        assert Z32.dtype == np.float32 and np.all(np.abs(Z32 - np.array(g["Z"])) <= 1e-3) and np.all(np.abs(r32 - np.array(g["reff"])) <= 1e-6)


def test_liu_hallett_and_const(built, orc):
    CMP = built.CMP
    g = G["liu_hallett"]
    one = lambda v: np.array([v])
    r = orc.diag_reff_lh97(g["rho_w"], one(g["rho"]), one(g["q_lcl"]), one(g["N_lcl"]), one(g["q_rai"]), one(g["N_rai"]))[0]
    assert abs(r - g["reff"]) <= g["atol"]
    # the three-argument method is N_lcl = 100, no rain (CloudDiagnostics.jl:150-165)
    assert orc.diag_reff_lh97(g["rho_w"], one(1.0), one(g["q_lcl"]))[0] == orc.diag_reff_lh97(g["rho_w"], one(1.0), one(g["q_lcl"]), one(100.0), one(0.0), one(0.0))[0]
    assert orc.diag_reff_lh97(g["rho_w"], one(1.0), one(1e-3), one(0.0), one(0.0), one(0.0))[0] == 0.0
    mp = CMP.Microphysics1MParams(np.float64)
    assert mp.block.cloud_liquid.r_eff == G["const"]["cloud_liquid"] and mp.block.cloud_ice.r_eff == G["const"]["cloud_ice"]
