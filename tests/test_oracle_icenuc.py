"""Pins the oracle's ice-nucleation / water-activity restatement on the reference's golden
values and cross-checks the ARG2000 restatement (for which the reference has no literal
goldens) against an independent 40-digit mpmath evaluation.  CPU only."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "icenuc_goldens.json")))
one = lambda v: np.array([v], dtype=np.float64)


def test_ice_nucleation_goldens(built, orc):
    CMP = built.CMP
    tps = CMP.ThermodynamicsParameters(np.float64)
    for name, da, J in G["deposition_J"]:
        got = orc.icenuc(CMP.pack_icenuc(tps, dust=CMP.DustType(name)), "deposition_J", one(da))[0][0]
        assert abs(got / J - 1) < 1e-12, (name, got, J)
    for name, da, J in G["ABIFM_J"]:
        got = orc.icenuc(CMP.pack_icenuc(tps, dust=CMP.DustType(name)), "ABIFM_J", one(da))[0][0]
        assert abs(got / J - 1) < 1e-9, (name, got, J)        # literals carry 14 digits
    blk = CMP.pack_icenuc(tps)
    k = G["koop"]
    assert abs(orc.icenuc(blk, "homogeneous_J_cubic", one(k["da_w"]))[0][0] / k["cubic"] - 1) < 1e-12
    assert abs(orc.icenuc(blk, "homogeneous_J_linear", one(k["da_w"]))[0][0] / k["linear"] - 1) < 1e-12
    out, nerr = orc.icenuc(blk, "homogeneous_J_cubic", np.array([0.1, 0.3, 0.5]))
    assert nerr == 2 and np.isnan(out[0]) and np.isfinite(out[1]) and np.isnan(out[2])   # DomainError outside [0.26, 0.34]
    assert abs(orc.icenuc(blk, "P3_deposition_N_i", one(240.0))[0][0] / G["P3_deposition_N_i"]["value"] - 1) < 1e-12
    assert orc.icenuc(blk, "P3_deposition_N_i", one(274.0))[0][0] == 0.0
    d = G["dust_fraction"]
    for name in ("DesertDust", "ArizonaTestDust"):
        got = orc.icenuc(CMP.pack_icenuc(tps, dust=CMP.DustType(name)), "dust_activated_number_fraction", one(d["Si"]), one(d["T"]))[0][0]
        assert abs(got / d[name] - 1) < 1e-8, (name, got)
    # unsupported aerosol type -> zero (IN:102, 134)
    assert orc.icenuc(CMP.pack_icenuc(tps, dust=CMP.DustType("Feldspar")), "ABIFM_J", one(0.2))[0][0] == 0.0
    assert orc.icenuc(CMP.pack_icenuc(tps, dust=CMP.DustType("DesertDust")), "deposition_J", one(0.2))[0][0] == 0.0


def test_water_activity_and_inpc_goldens(built, orc):
    CMP = built.CMP
    tps = CMP.ThermodynamicsParameters(np.float64)
    blk = CMP.pack_icenuc(tps)
    g = G["a_w_ice"]
    assert abs(orc.icenuc(blk, "a_w_ice", one(g["T"]))[0][0] / g["value"] - 1) < 1e-14
    g = G["a_w_eT"]
    assert abs(orc.icenuc(blk, "a_w_eT", one(g["T"]), one(g["e"]))[0][0] / g["value"] - 1) < 1e-14
    g = G["h2so4"]
    assert abs(orc.icenuc(blk, "H2SO4_soln_saturation_vapor_pressure", one(g["T"]), one(g["x"]))[0][0] / g["p_sol"] - 1) < 1e-12
    assert abs(orc.icenuc(blk, "a_w_xT", one(g["T"]), one(g["x"]))[0][0] / g["a_w"] - 1) < 1e-13
    g = G["frostenberg_mean"]
    T = 273.15 + g["T_celsius"]
    assert abs(orc.icenuc(blk, "INP_concentration_mean", one(T))[0][0] / g["value"] - 1) < 1e-14
    b2 = CMP.pack_icenuc(tps, frostenberg=CMP.FrostenbergParameters(np.float64, overrides={"Frostenberg2023_a_coefficient": 2.0}))
    assert abs(orc.icenuc(b2, "INP_concentration_mean", one(T))[0][0] / g["a2"] - 1) < 1e-14
    b3 = CMP.pack_icenuc(tps, frostenberg=CMP.FrostenbergParameters(np.float64, overrides={"Frostenberg2023_b_coefficient": 2.0}))
    assert abs(orc.icenuc(b3, "INP_concentration_mean", one(T))[0][0] / g["b2"] - 1) < 1e-14
    assert np.isneginf(orc.icenuc(blk, "INP_concentration_mean", one(280.0))[0][0])   # log(0) = -Inf above freezing


def _mp_arg(blk, T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice):
    """Independent evaluation of AA:138-259 with mpmath (40 digits)."""
    import mpmath as mp
    mp.mp.dps = 40
    f = lambda x: mp.mpf(float(x))
    t, ap, aps = blk.tps, blk.arg, blk.aps
    T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice = map(f, (T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice))
    Rv, Rd = f(t.R_v), f(t.R_d)
    psat = lambda LH0, dcp: f(t.press_triple) * (T / f(t.T_triple)) ** (dcp / Rv) * mp.exp((LH0 - dcp * f(t.T_0)) / Rv * (1 / f(t.T_triple) - 1 / T))
    pvs, pvi = psat(f(t.LH_v0), f(t.cp_v) - f(t.cp_l)), psat(f(t.LH_s0), f(t.cp_v) - f(t.cp_i))
    Lv = f(t.LH_v0) + (f(t.cp_v) - f(t.cp_l)) * (T - f(t.T_0))
    Ls = f(t.LH_s0) + (f(t.cp_v) - f(t.cp_i)) * (T - f(t.T_0))
    Rm = Rd * (1 + (Rv / Rd - 1) * q_tot - Rv / Rd * (q_liq + q_ice))
    cpm = f(t.cp_d) + (f(t.cp_v) - f(t.cp_d)) * q_tot + (f(t.cp_l) - f(t.cp_v)) * q_liq + (f(t.cp_i) - f(t.cp_v)) * q_ice
    rho = p / (Rm * T)
    pv = (q_tot - q_liq - q_ice) * rho * Rv * T
    Gf = lambda L, ps: 1 / (L / f(aps.K_therm) / T * (L / Rv / T - 1) + Rv * T / f(aps.D_vapor) / ps)
    G = Gf(Lv, pvs) / f(ap.rho_w)
    g_ = f(ap.g)
    alpha = pv / pvs * (Lv * g_ / Rv / cpm / T ** 2 - g_ / Rm / T)
    gamma = Rv * T / pvs + pv / pvs * Rm * Lv ** 2 / Rv / cpm / T / p
    A = 2 * f(ap.sigma) * f(ap.M_w) / f(ap.rho_w) / f(ap.R) / T
    zeta = 2 * A / 3 * mp.sqrt(alpha * w / G)
    tmp, Sm = mp.mpf(0), []
    for i in range(blk.n_modes):
        m = blk.modes[i]
        sm = 2 / mp.sqrt(f(m.hygro)) * (A / 3 / f(m.r_dry)) ** mp.mpf(1.5)
        Sm.append(sm)
        ls = mp.log(f(m.stdev))
        ff = f(ap.f1) * mp.exp(f(ap.f2) * ls ** 2)
        gg = f(ap.g1) + f(ap.g2) * ls
        eta = mp.sqrt(alpha * w / G) ** 3 / (2 * mp.pi * f(ap.rho_w) * gamma * f(m.N))
        tmp += 1 / sm ** 2 * (ff * (zeta / eta) ** f(ap.p1) + gg * (sm ** 2 / (eta + 3 * zeta)) ** f(ap.p2))
    S_arg = 1 / mp.sqrt(tmp)
    r_liq = mp.mpf(0) if N_liq < 2.2e-16 else mp.cbrt(rho * q_liq / N_liq / f(ap.rho_w) / (mp.mpf(4) / 3 * mp.pi))
    K_liq = 4 * mp.pi * f(ap.rho_w) * N_liq * r_liq * G * gamma
    gamma_i = Rv * T / pvs + pv / pvs * Rm * Lv * Ls / Rv / cpm / T / p
    r_ice = mp.mpf(0) if N_ice < 2.2e-16 else mp.cbrt(rho * q_ice / N_ice / f(ap.rho_i) / (mp.mpf(4) / 3 * mp.pi))
    K_ice = 4 * mp.pi * N_ice * r_ice * Gf(Ls, pvi) * gamma_i
    xi = pvs / pvi
    S = max(mp.mpf(0), S_arg * (alpha * w - K_ice * (xi - 1)) / (alpha * w + (K_liq + K_ice * xi) * S_arg))
    N = []
    for i in range(blk.n_modes):
        m = blk.modes[i]
        if S == 0:
            N.append(mp.mpf(0))
        else:
            u = 2 * mp.log(Sm[i] / S) / 3 / mp.sqrt(2) / mp.log(f(m.stdev))
            N.append(f(m.N) * (1 - mp.erf(u)) / 2)
    return S, N


@pytest.mark.parametrize("kind,hyd", [("kappa", False), ("B", True)])
def test_arg2000_against_mpmath(built, orc, kind, hyd):
    CMP = built.CMP
    from cumicro.testing import synthetic_states_activation, arg_test_distribution
    tps = CMP.ThermodynamicsParameters(np.float64)
    blk = CMP.pack_icenuc(tps, ad=arg_test_distribution(kind))
    assert blk.n_modes == 3
    st = synthetic_states_activation(40, seed=5, with_hydrometeors=hyd)
    keys = ("T", "p", "w", "q_tot", "q_liq", "q_ice", "N_liq", "N_ice")
    out = orc.arg_icenuc(blk, *[st[k] for k in keys])
    for i in range(40):
        S, N = _mp_arg(blk, *[st[k][i] for k in keys])
        assert abs(out["S_max"][i] - float(S)) <= 1e-12 * float(S) + 1e-300, (i, out["S_max"][i], float(S))
        for m in range(3):
            # 1 - erf(u) cancels for u << 0 (fully activated mode): compare on the scale of N
            assert abs(out["N_act"][m][i] - float(N[m])) <= 1e-12 * blk.modes[m].N, (i, m)
    assert (out["S_max"] > 0).any()


def test_kappa_and_B_descriptions_agree(built, orc):
    """test/gpu_tests.jl:580-587: with kappa_j = B_j the two mode types give the same answer (rtol 1e-5
    there because of the tabulated kappas; here B is converted exactly)."""
    from cumicro import AerosolModel as AM, AerosolActivation as AA
    CMP = built.CMP
    ap = CMP.AerosolActivationParameters(np.float64)
    B = AM.AerosolDistribution((AM.Mode_B(0.05e-6, 2.0, 1e8, (1.0,), (1.0,), (1.0,), (0.132,), (3.0,), (1770.0,)),))
    hB = AA.mean_hygroscopicity_parameter(ap, B)[0]
    K = AM.AerosolDistribution((AM.Mode_κ(0.05e-6, 2.0, 1e8, (1.0,), (1.0,), (0.132,), (float(hB),)),))
    assert abs(AA.mean_hygroscopicity_parameter(ap, K)[0] / hB - 1) < 1e-15
    assert abs(hB / (3.0 * 1.0 * 1.0 / 0.132 * 1770.0 * 0.01801528 / 1000.0) - 1) < 1e-12


def test_multi_argument_rates_goldens(built, orc):
    """MohlerDepositionRate, P3_het_N_i, INP_concentration_frequency, P3.het_ice_nucleation on the reference's literals."""
    CMP = built.CMP
    tps = CMP.ThermodynamicsParameters(np.float64)
    g = G["mohler_rate"]
    for name in ("DesertDust", "ArizonaTestDust"):
        blk = CMP.pack_icenuc(tps, dust=CMP.DustType(name))
        out, _, nerr = orc.icenuc_rates(blk, "MohlerDepositionRate", one(g["Si"]), one(g["T"]), one(g["dSi_dt"]), one(g["N_aer"]))
        assert nerr == 0 and abs(out[0] / g[name] - 1) < 1e-9
    out, _, nerr = orc.icenuc_rates(CMP.pack_icenuc(tps, dust=CMP.DustType("DesertDust")), "MohlerDepositionRate", one(1.4), one(240.0), one(0.03), one(3000.0))
    assert nerr == 1 and np.isnan(out[0])                      # @assert Si < Sᵢ_max  (IN:73)
    g = G["P3_het_N_i"]
    out, _, _ = orc.icenuc_rates(CMP.pack_icenuc(tps), "P3_het_N_i", one(g["T"]), one(g["N_l"]), one(g["V_l"]), one(g["dt"]))
    assert abs(out[0] / g["value"] - 1) < 1e-12
    out, _, _ = orc.icenuc_rates(CMP.pack_icenuc(tps), "INP_concentration_frequency", one(220000.0), one(233.0))
    assert abs(out[0] / 0.26 - 1) < 0.1                        # test/gpu_tests.jl:1050 (rtol 0.1; pins sigma loosely)
    assert orc.icenuc_rates(CMP.pack_icenuc(tps), "INP_concentration_frequency", one(220000.0), one(274.0))[0][0] == 0.0
    # P3.het_ice_nucleation sweep (test/p3_tests.jl:573-614)
    g = G["p3_het_freezing"]
    d = CMP.DEFAULTS
    Rd, Rv = d["gas_constant_dry_air"], d["gas_constant_vapor"]
    from cumicro.testing import psat_liq
    blk = CMP.pack_icenuc(tps, dust=CMP.DustType(g["aerosol"]))
    for qv, rN, rL in zip(g["qv"], g["dNdt"], g["dLdt"]):
        eps_ = Rd / Rv
        e_v = g["p"] * qv / (eps_ + qv * (1 - eps_))
        RH = e_v / psat_liq(g["T"])
        q_tot = qv + g["q_lcl"]
        R_m = Rd * (1 + (Rv / Rd - 1) * q_tot - Rv / Rd * g["q_lcl"])
        rho = g["p"] / (R_m * g["T"])
        dN, dL, _ = orc.icenuc_rates(blk, "het_ice_nucleation", one(g["q_lcl"]), one(g["N_lcl"]), one(RH), one(g["T"]), one(rho))
        assert abs(dN[0] / rN - 1) < g["rtol"] and abs(dL[0] / rL - 1) < g["rtol"], (qv, dN[0], rN, dL[0], rL)
