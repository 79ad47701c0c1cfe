"""GPU parity of the fused 1M + 2M + ice-nucleation kernel (BASELINE config 5): every output
column equals what the separate kernels / the oracle give, and the in-kernel diagnostics equal
the Float64 sums of the per-point contributions."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fused_matches_the_separate_paths_and_the_oracle(built, orc, cuda):
    import torch
    from cumicro.testing import synthetic_states_fused, arg_test_distribution, assert_parity
    from cumicro import fused
    CMP, BMT, AA = built.CMP, built.BMT, built.AA
    n = (1 << 17) + 3
    st = synthetic_states_fused(n, seed=5)
    d = {k: torch.from_numpy(v).to(cuda) for k, v in st.items()}
    tps = CMP.ThermodynamicsParameters(np.float64)
    mp1, mp2 = CMP.Microphysics1MParams(np.float64), CMP.Microphysics2MParams(np.float64)
    ad = arg_test_distribution("kappa")
    blk3 = CMP.pack_icenuc(tps, ad=ad, dust=CMP.DustType("Kaolinite"), hom_linear=True)
    out = fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *[d[k] for k in fused.IN_NAMES])
    # (1) bit-identical to the separate kernels (same device functions, same arithmetic)
    o1 = BMT.bulk_microphysics_tendencies(BMT.Instantaneous(), BMT.Microphysics1Moment(), mp1, tps, d["rho"], d["T"], d["q_tot"],
                                          d["q_lcl"], d["q_icl"], d["q_rai"], d["q_sno"])
    for a, b in zip(("m1_dq_lcl_dt", "m1_dq_icl_dt", "m1_dq_rai_dt", "m1_dq_sno_dt"), ("dq_lcl_dt", "dq_icl_dt", "dq_rai_dt", "dq_sno_dt")):
        assert torch.equal(out[a], o1[b]), a
    o2 = BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp2, tps, d["rho"], d["T"], d["q_tot"], d["q_lcl"], d["n_lcl"],
                                          d["q_rai"], d["n_rai"], q_ice=d["q_icl"] + d["q_sno"])
    for a, b in zip(("m2_dq_lcl_dt", "m2_dn_lcl_dt", "m2_dq_rai_dt", "m2_dn_rai_dt"), ("dq_lcl_dt", "dn_lcl_dt", "dq_rai_dt", "dn_rai_dt")):
        assert torch.equal(out[a], o2[b]), a
    # (2) ice nucleation / activation against the oracle
    args = (st["T"], st["p"], st["w"], st["q_tot"], st["q_lcl"] + st["q_rai"], st["q_icl"] + st["q_sno"], st["rho"] * st["n_lcl"], np.zeros(n))
    ref = orc.arg_icenuc(blk3, *args)
    bnd = orc.arg_icenuc(blk3, *args, bound=True)
    for a in ("J_dep", "J_ABIFM", "J_hom"):
        rep = assert_parity(a, out[a].cpu().numpy(), ref[a], bound=bnd[a])
        assert rep["max_rel"] <= 1e-12, (a, rep)
    # (3) diagnostics = Float64 sums of the per-point contributions
    diag = out["diag"].cpu().numpy()
    rho = st["rho"]
    c0 = rho * (out["m1_dq_rai_dt"].cpu().numpy() + out["m1_dq_sno_dt"].cpu().numpy())
    c1 = rho * out["m2_dq_rai_dt"].cpu().numpy()
    c2 = ref["N_act"][0] + ref["N_act"][1] + ref["N_act"][2]
    for got, contrib in ((diag[0], c0), (diag[1], c1), (diag[2], c2)):
        assert abs(got - contrib.sum()) <= 1e-11 * np.abs(contrib).sum()
    assert diag[3] == n
    # bit-reproducible reduction
    again = fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *[d[k] for k in fused.IN_NAMES])
    assert torch.equal(again["diag"], out["diag"])
    # slabs: the diagnostics of two half slabs add up to the whole (what the all-reduce does across GPUs)
    lo, hi = fused.slab_bounds(n, 2, 0)
    s0 = fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *[d[k][lo:hi].contiguous() for k in fused.IN_NAMES])
    lo1, hi1 = fused.slab_bounds(n, 2, 1)
    s1 = fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *[d[k][lo1:hi1].contiguous() for k in fused.IN_NAMES])
    np.testing.assert_allclose((s0["diag"] + s1["diag"]).cpu().numpy(), diag, rtol=1e-12)
    assert torch.equal(s1["m2_dn_rai_dt"], out["m2_dn_rai_dt"][lo1:hi1])


def test_fused_at_the_full_config5_size_2pow24(built, cuda):
    """BASELINE config 5 runs 2^24 points per GPU: at that size every tendency column of the fused kernel is bit-identical to the
    separately launched 1-moment and 2-moment kernels (each of which is checked against the oracle at its own full size), the
    nucleation rates to the stand-alone ARG2000 / ice-nucleation kernel within its erf variant's rounding, and the diagnostics to
    the Float64 sums of the columns."""
    import torch
    from cumicro.testing import synthetic_states_fused, arg_test_distribution
    from cumicro import fused
    CMP, BMT, AA = built.CMP, built.BMT, built.AA
    n = 1 << 24
    st = synthetic_states_fused(n, seed=2024)
    d = {k: torch.from_numpy(v).to(cuda) for k, v in st.items()}
    tps = CMP.ThermodynamicsParameters(np.float64)
    mp1, mp2 = CMP.Microphysics1MParams(np.float64), CMP.Microphysics2MParams(np.float64)
    ad = arg_test_distribution("kappa")
    blk3 = CMP.pack_icenuc(tps, ad=ad, dust=CMP.DustType("Kaolinite"), hom_linear=True)
    out = fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *[d[k] for k in fused.IN_NAMES])
    o1 = BMT.bulk_microphysics_tendencies(BMT.Instantaneous(), BMT.Microphysics1Moment(), mp1, tps, d["rho"], d["T"], d["q_tot"],
                                          d["q_lcl"], d["q_icl"], d["q_rai"], d["q_sno"])
    for a, b in zip(("m1_dq_lcl_dt", "m1_dq_icl_dt", "m1_dq_rai_dt", "m1_dq_sno_dt"), ("dq_lcl_dt", "dq_icl_dt", "dq_rai_dt", "dq_sno_dt")):
        assert torch.equal(out[a], o1[b]), a
    del o1
    o2 = BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp2, tps, d["rho"], d["T"], d["q_tot"], d["q_lcl"], d["n_lcl"],
                                          d["q_rai"], d["n_rai"], q_ice=d["q_icl"] + d["q_sno"])
    for a, b in zip(("m2_dq_lcl_dt", "m2_dn_lcl_dt", "m2_dq_rai_dt", "m2_dn_rai_dt"), ("dq_lcl_dt", "dn_lcl_dt", "dq_rai_dt", "dn_rai_dt")):
        assert torch.equal(out[a], o2[b]), a
    del o2
    sep = AA.activation_and_ice_nucleation(CMP.AerosolActivationParameters(np.float64), ad, CMP.AirProperties(np.float64), tps,
                                           CMP.DustType("Kaolinite"), CMP.Koop2000(np.float64), d["T"], d["p"], d["w"], d["q_tot"],
                                           d["q_lcl"] + d["q_rai"], d["q_icl"] + d["q_sno"], d["rho"] * d["n_lcl"], torch.zeros_like(d["T"]),
                                           hom_linear=True)
    for a in ("J_dep", "J_ABIFM", "J_hom"):
        g, r = out[a], sep[a]
        ok = torch.isfinite(r)
        assert torch.equal(torch.isfinite(g), ok), a
        assert float(((g[ok] - r[ok]).abs() / r[ok].abs().clamp_min(1e-300)).max()) <= 1e-12, a
    diag = out["diag"].cpu().numpy()
    c0 = (d["rho"] * (out["m1_dq_rai_dt"] + out["m1_dq_sno_dt"])).sum(dtype=torch.float64).item()
    c1 = (d["rho"] * out["m2_dq_rai_dt"]).sum(dtype=torch.float64).item()
    a0 = (d["rho"] * (out["m1_dq_rai_dt"] + out["m1_dq_sno_dt"])).abs().sum(dtype=torch.float64).item()
    a1 = (d["rho"] * out["m2_dq_rai_dt"]).abs().sum(dtype=torch.float64).item()
    assert abs(diag[0] - c0) <= 1e-10 * a0 and abs(diag[1] - c1) <= 1e-10 * a1
    nact = sep["N_act"][0] + sep["N_act"][1] + sep["N_act"][2]
    nact = torch.where((d["w"] > 0) & torch.isfinite(nact), nact, torch.zeros_like(nact)).sum(dtype=torch.float64).item()
    assert abs(diag[2] - nact) <= 1e-10 * abs(nact)
    assert diag[3] == n
