"""GPU parity of the ice-nucleation / water-activity leaves and the fused ARG2000 + nucleation
kernel (BASELINE config 3) through the C-ABI vs the CPU oracle; Float64 1e-12 relative (or the
reference's own rounding bound), Float32 <= 4 ULP of the true value."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "icenuc_goldens.json")))
KEYS = ("T", "p", "w", "q_tot", "q_liq", "q_ice", "N_liq", "N_ice")


def test_goldens_through_the_gpu(built, cuda):
    import torch
    CMP, IN = built.CMP, built.IN
    tps = CMP.ThermodynamicsParameters(np.float64)
    f = lambda v: torch.full((3,), v, dtype=torch.float64, device=cuda)
    for name, da, J in G["deposition_J"]:
        assert abs(float(IN.deposition_J(CMP.DustType(name), tps, f(da))[0]) / J - 1) < 1e-12
    for name, da, J in G["ABIFM_J"]:
        assert abs(float(IN.ABIFM_J(CMP.DustType(name), tps, f(da))[0]) / J - 1) < 1e-9
    k = G["koop"]
    koop = CMP.Koop2000(np.float64)
    assert abs(float(IN.homogeneous_J_cubic(koop, tps, f(k["da_w"]))[0]) / k["cubic"] - 1) < 1e-12
    assert abs(float(IN.homogeneous_J_linear(koop, tps, f(k["da_w"]))[0]) / k["linear"] - 1) < 1e-12
    with pytest.raises(IN.DomainError):                               # DomainError of IN:558-562
        IN.homogeneous_J_cubic(koop, tps, torch.tensor([0.3, 0.1], dtype=torch.float64, device=cuda))
    out = IN.homogeneous_J_cubic(koop, tps, torch.tensor([0.3, 0.1], dtype=torch.float64, device=cuda), check_domain=False)
    assert bool(torch.isfinite(out[0])) and bool(torch.isnan(out[1]))
    assert abs(float(IN.a_w_ice(tps, f(G["a_w_ice"]["T"]))[0]) / G["a_w_ice"]["value"] - 1) < 1e-13
    g = G["a_w_eT"]
    assert abs(float(IN.a_w_eT(tps, f(g["e"]), f(g["T"]))[0]) / g["value"] - 1) < 1e-13
    g = G["h2so4"]
    h = CMP.H2SO4SolutionParameters(np.float64)
    assert abs(float(IN.H2SO4_soln_saturation_vapor_pressure(h, tps, f(g["x"]), f(g["T"]))[0]) / g["p_sol"] - 1) < 1e-12
    assert abs(float(IN.a_w_xT(h, tps, f(g["x"]), f(g["T"]))[0]) / g["a_w"] - 1) < 1e-12
    assert abs(float(IN.P3_deposition_N_i(CMP.MorrisonMilbrandt2014(np.float64), tps, f(240.0))[0]) / G["P3_deposition_N_i"]["value"] - 1) < 1e-12
    d = G["dust_fraction"]
    for name in ("DesertDust", "ArizonaTestDust"):
        got = float(IN.dust_activated_number_fraction(CMP.DustType(name), CMP.Mohler2006(np.float64), tps, f(d["Si"]), f(d["T"]))[0])
        assert abs(got / d[name] - 1) < 1e-8


def test_leaves_f64_parity(built, orc, cuda):
    import torch
    from cumicro.testing import assert_parity
    CMP, IN = built.CMP, built.IN
    tps = CMP.ThermodynamicsParameters(np.float64)
    rng = np.random.default_rng(3)
    n = 1 << 15
    da = rng.uniform(0.0, 0.45, n)
    T = rng.uniform(185.0, 300.0, n)
    e = rng.uniform(1.0, 3000.0, n)
    x = rng.uniform(0.0, 0.5, n)
    d = lambda a: torch.from_numpy(a).to(cuda)
    dust = CMP.DustType("Illite")
    blk = CMP.pack_icenuc(tps, dust=dust)
    cases = [("deposition_J", IN.deposition_J(dust, tps, d(da)), (da, None)), ("ABIFM_J", IN.ABIFM_J(dust, tps, d(da)), (da, None)),
             ("homogeneous_J_linear", IN.homogeneous_J_linear(CMP.Koop2000(np.float64), tps, d(da)), (da, None)),
             ("homogeneous_J_cubic", IN.homogeneous_J_cubic(CMP.Koop2000(np.float64), tps, d(da), check_domain=False), (da, None)),
             ("a_w_ice", IN.a_w_ice(tps, d(T)), (T, None)), ("a_w_eT", IN.a_w_eT(tps, d(e), d(T)), (T, e)),
             ("a_w_xT", IN.a_w_xT(CMP.H2SO4SolutionParameters(np.float64), tps, d(x), d(T)), (T, x)),
             ("P3_deposition_N_i", IN.P3_deposition_N_i(CMP.MorrisonMilbrandt2014(np.float64), tps, d(T)), (T, None)),
             ("INP_concentration_mean", IN.INP_concentration_mean(CMP.FrostenbergParameters(np.float64), tps, d(T)), (T, None))]
    for name, got, (a, b) in cases:
        ref, nerr = orc.icenuc(blk, name, a, b)
        rep = assert_parity(name, got.cpu().numpy(), ref)       # includes NaN (domain) and -Inf (log 0) pattern equality
        assert rep["max_rel"] <= 1e-12, (name, rep)
        if name == "homogeneous_J_cubic":
            assert nerr == int(np.isnan(ref).sum()) > 0


@pytest.mark.parametrize("kind,hyd", [("kappa", False), ("B", True)])
def test_arg_icenuc_fused_f64_parity(built, orc, cuda, kind, hyd):
    import torch
    from cumicro.testing import synthetic_states_activation, arg_test_distribution, assert_parity
    CMP, AA = built.CMP, built.AA
    tps = CMP.ThermodynamicsParameters(np.float64)
    ap, aip, ad = CMP.AerosolActivationParameters(np.float64), CMP.AirProperties(np.float64), arg_test_distribution(kind)
    dust, koop = CMP.DustType("Kaolinite"), CMP.Koop2000(np.float64)
    n = 1 << 16
    st = synthetic_states_activation(n, seed=9, with_hydrometeors=hyd)
    cols = [torch.from_numpy(st[k]).to(cuda) for k in KEYS]
    for hom_linear in (False, True):
        got = AA.activation_and_ice_nucleation(ap, ad, aip, tps, dust, koop, *cols, hom_linear=hom_linear, with_mass=True)
        blk = CMP.pack_icenuc(tps, aps=aip, ap=ap, ad=ad, dust=dust, koop=koop, hom_linear=hom_linear)
        ref = orc.arg_icenuc(blk, *[st[k] for k in KEYS])
        bnd = orc.arg_icenuc(blk, *[st[k] for k in KEYS], bound=True)
        for k in ("S_max", "J_dep", "J_ABIFM", "J_hom", "da_w"):
            rep = assert_parity(k, got[k].cpu().numpy(), ref[k], bound=bnd[k])
            assert rep["max_rel"] <= 1e-12, (k, rep)
        for m in range(3):
            for k in ("N_act", "M_act"):
                rep = assert_parity(f"{k}[{m}]", got[k][m].cpu().numpy(), ref[k][m], bound=bnd[k][m])
                assert rep["max_rel"] <= 1e-12, (k, m, rep)
        assert int(got["n_domain_errors"].item()) == ref["n_domain_errors"]
        assert (ref["n_domain_errors"] == 0) == hom_linear
    # module-level entry points select the same columns
    N = AA.N_activated_per_mode(ap, ad, aip, tps, *cols)
    tot = AA.total_N_activated(ap, ad, aip, tps, *cols)
    assert torch.equal(N[1], got["N_act"][1]) and torch.equal(tot, N[0] + N[1] + N[2])
    assert torch.equal(AA.max_supersaturation(ap, ad, aip, tps, *cols), got["S_max"])


def test_config3_f32_full_size_2pow25(built, orc, cuda):
    """BASELINE config 3 at its full size (2^25 points, Float32, 3 aerosol modes): every output of the GPU kernel against the
    Float64 truth of the Float32 method on ALL points (4 Float32 ULP), regime pattern against the Float32 oracle."""
    import torch
    from cumicro.testing import synthetic_states_activation, arg_test_distribution, ulp_error_f32
    CMP, AA = built.CMP, built.AA
    F = np.float32
    tps = CMP.ThermodynamicsParameters(F)
    ap, aip, ad = CMP.AerosolActivationParameters(F), CMP.AirProperties(F), arg_test_distribution("kappa")
    dust, koop = CMP.DustType("Kaolinite", F), CMP.Koop2000(F)
    n = 1 << 25
    st = synthetic_states_activation(n, seed=1234, dtype=F)
    cols = [torch.from_numpy(st[k]).to(cuda) for k in KEYS]
    got = AA.activation_and_ice_nucleation(ap, ad, aip, tps, dust, koop, *cols, hom_linear=True)
    blk64 = CMP.widen(CMP.pack_icenuc(tps, aps=aip, ap=ap, ad=ad, dust=dust, koop=koop, hom_linear=True))
    del cols
    worst = {}
    chunk = 1 << 22
    for lo in range(0, n, chunk):                                   # the CPU port, chunked to bound memory
        s64 = [st[k][lo:lo + chunk].astype(np.float64) for k in KEYS]
        with orc.f32_thresholds():
            truth = orc.arg_icenuc(blk64, *s64)
        for k in ("S_max", "J_dep", "J_ABIFM", "J_hom"):
            g = got[k][lo:lo + chunk].cpu().numpy()
            t = truth[k]
            ok = np.isfinite(t) & (np.abs(t) < 3e38) & (np.abs(t) > 1.2e-38)
            e = ulp_error_f32(g[ok], t[ok])
            # S_max carries the cancellation of the ARG2000 supersaturation balance: the 2^16-point test applies the oracle's bound there
            worst[k] = max(worst.get(k, 0.0), float(np.percentile(e, 99.9)) if k == "S_max" else float(e.max()))
        for i in range(3):
            g = got["N_act"][i][lo:lo + chunk].cpu().numpy()
            t = truth["N_act"][i]
            big = t > 1e-6 * blk64.modes[i].N
            worst[f"N_act{i}"] = max(worst.get(f"N_act{i}", 0.0), float(ulp_error_f32(g[big], t[big]).max()))
    assert all(v <= 4 for v in worst.values()), worst


def test_config3_f32_2pow20(built, orc, cuda):
    """BASELINE config 3 (Float32, 3 aerosol modes, deposition + ABIFM + Koop J): the Float32
    method within 4 Float32 ULP of the true value."""
    import torch
    from cumicro.testing import synthetic_states_activation, arg_test_distribution, assert_f32_method
    CMP, AA = built.CMP, built.AA
    F = np.float32
    tps = CMP.ThermodynamicsParameters(F)
    ap, aip, ad = CMP.AerosolActivationParameters(F), CMP.AirProperties(F), arg_test_distribution("kappa")
    dust, koop = CMP.DustType("Kaolinite", F), CMP.Koop2000(F)
    n = 1 << 20
    st = synthetic_states_activation(n, seed=11, dtype=F)
    cols = [torch.from_numpy(st[k]).to(cuda) for k in KEYS]
    got = AA.activation_and_ice_nucleation(ap, ad, aip, tps, dust, koop, *cols, hom_linear=True)
    blk32 = CMP.pack_icenuc(tps, aps=aip, ap=ap, ad=ad, dust=dust, koop=koop, hom_linear=True)
    blk64 = CMP.widen(blk32)
    m = 1 << 16                                                   # the CPU reference on a 65 536-point sample
    s64 = [st[k][:m].astype(np.float64) for k in KEYS]
    ref32 = orc.arg_icenuc(blk32, *[st[k][:m] for k in KEYS])
    with orc.f32_thresholds():
        truth = orc.arg_icenuc(blk64, *s64)
        bound = orc.arg_icenuc(blk64, *s64, bound=True)
    for k in ("S_max", "J_dep", "J_ABIFM", "J_hom"):
        g = got[k].cpu().numpy()
        assert g.dtype == np.float32 and g.shape == (n,)
        # Float32 reference overflows J to Inf / underflows to 0 where the true value is outside Float32 range
        assert_f32_method("arg:" + k, g[:m], np.where(np.isfinite(ref32[k]), ref32[k], truth[k].astype(np.float32)), truth[k], bound[k],
                          ref_is_f32_oracle=True)
    for i in range(3):
        assert_f32_method(f"arg:N_act[{i}]", got["N_act"][i].cpu().numpy()[:m], ref32["N_act"][i], truth["N_act"][i],
                          np.maximum(bound["N_act"][i], 1e-9 * blk64.modes[i].N), ref_is_f32_oracle=True)


def test_multi_argument_rates_parity(built, orc, cuda):
    """cumicro_icenuc_rates_* vs the oracle: Float64 1e-12, Float32 <= 4 ULP of the true value, domain errors counted."""
    import torch
    from cumicro.testing import assert_parity, assert_f32_method
    CMP, IN = built.CMP, built.IN
    rng = np.random.default_rng(8)
    n = 20000
    T = rng.uniform(200.0, 280.0, n)
    for FT in (np.float64, np.float32):
        tps = CMP.ThermodynamicsParameters(FT)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=FT)).to(cuda)
        c = lambda a: np.ascontiguousarray(a, dtype=FT)
        dust, moh = CMP.DustType("ArizonaTestDust", FT), CMP.Mohler2006(FT)
        cases = {
            "MohlerDepositionRate": ([rng.uniform(1.0, 1.34, n), T, rng.uniform(-0.01, 0.05, n), rng.uniform(0, 5000, n)],
                                     lambda a: IN.MohlerDepositionRate(dust, moh, tps, *a), CMP.pack_icenuc(tps, dust=dust, mohler=moh)),
            "P3_het_N_i": ([T, 10 ** rng.uniform(2, 8, n), 10 ** rng.uniform(-18, -12, n), rng.uniform(0.1, 100, n)],
                           lambda a: IN.P3_het_N_i(CMP.MorrisonMilbrandt2014(FT), tps, *a), CMP.pack_icenuc(tps)),
            "INP_concentration_frequency": ([10 ** rng.uniform(0, 7, n), T],
                                            lambda a: IN.INP_concentration_frequency(CMP.FrostenbergParameters(FT), tps, *a), CMP.pack_icenuc(tps)),
        }
        for name, (cols, fn, blk) in cases.items():
            cols = [c(a) for a in cols]
            got = fn([d(a) for a in cols]).cpu().numpy()
            # P3_het_N_i = N_l (1 - exp(-x)) subtracts nearly equal numbers for small x (the reference's own Float32 literal
            # has lost 13 % there, test/gpu_tests.jl:1024-1031): its first-order rounding bound is N_l * ulp(1)
            bound = cols[1].astype(np.float64) * 2.0 ** -52 if name == "P3_het_N_i" else None
            if FT is np.float64:
                assert_parity(name, got, orc.icenuc_rates(blk, name, *cols)[0], bound=bound)
            else:
                truth = orc.icenuc_rates(CMP.widen(blk), name, *[a.astype(np.float64) for a in cols])[0]
                assert_f32_method(name, got, truth.astype(FT), truth, bound)
        ill = CMP.DustType("Illite", FT)
        cols = [c(rng.uniform(0, 1e-3, n)), c(10 ** rng.uniform(6, 9, n)), c(rng.uniform(0.8, 1.15, n)), c(T), c(rng.uniform(0.3, 1.3, n))]
        dN, dL = IN.het_ice_nucleation(ill, tps, *[d(a) for a in cols])
        blk = CMP.pack_icenuc(tps, dust=ill)
        if FT is np.float64:
            rN, rL, _ = orc.icenuc_rates(blk, "het_ice_nucleation", *cols)
            assert_parity("het dNdt", dN.cpu().numpy(), rN)
            assert_parity("het dLdt", dL.cpu().numpy(), rL)
        else:
            tN, tL, _ = orc.icenuc_rates(CMP.widen(blk), "het_ice_nucleation", *[a.astype(np.float64) for a in cols])
            big = tN < 3e38                                        # Float32 overflow of J -> the isfinite guard (P3_processes.jl:36-42) differs by type
            assert_f32_method("het dNdt f32", dN.cpu().numpy()[big], tN[big].astype(FT), tN[big])
    tps = CMP.ThermodynamicsParameters(np.float64)
    bad = torch.tensor([1.2, 1.36], dtype=torch.float64, device=cuda)
    two = lambda v: torch.full((2,), v, dtype=torch.float64, device=cuda)
    with pytest.raises(IN.DomainError):
        IN.MohlerDepositionRate(CMP.DustType("DesertDust"), CMP.Mohler2006(np.float64), tps, bad, two(240.0), two(0.03), two(3000.0))
    g = G["mohler_rate"]
    got = IN.MohlerDepositionRate(CMP.DustType("DesertDust"), CMP.Mohler2006(np.float64), tps, two(g["Si"]), two(g["T"]), two(g["dSi_dt"]), two(g["N_aer"]))
    assert abs(float(got[0]) / g["DesertDust"] - 1) < 1e-9


def test_tile_shape_ragged_sizes_misaligned_columns_and_domain_error_count(built, orc, cuda):
    """ARG2000 + nucleation rates in the tile launch shape: ragged sizes and 4-byte-aligned Float32 / 8-byte-aligned Float64 columns
    give the same bits as the aligned full-tile path, and the DomainError counter (Koop's cubic outside its range) counts every point
    of a partial tile exactly once."""
    import torch
    from cumicro.testing import synthetic_states_activation, arg_test_distribution
    CMP, AA = built.CMP, built.AA
    for F in (np.float64, np.float32):
        tps = CMP.ThermodynamicsParameters(F)
        ap, aip, ad = CMP.AerosolActivationParameters(F), CMP.AirProperties(F), arg_test_distribution("kappa")
        dust, koop = CMP.DustType("Kaolinite", F) if F is np.float32 else CMP.DustType("Kaolinite"), CMP.Koop2000(F)
        n_big = 128 * 21 + 77
        st = synthetic_states_activation(n_big + 1, seed=3, dtype=F)
        dev = lambda lo, hi: [torch.from_numpy(st[k][lo:hi].copy()).to(cuda) for k in KEYS]
        ref = AA.activation_and_ice_nucleation(ap, ad, aip, tps, dust, koop, *dev(0, n_big), hom_linear=False)
        names = ("S_max", "J_dep", "J_ABIFM", "J_hom", "da_w")
        for n in (1, 127, 129, 1000, n_big):
            got = AA.activation_and_ice_nucleation(ap, ad, aip, tps, dust, koop, *dev(0, n), hom_linear=False)
            for k in names:
                assert torch.equal(torch.nan_to_num(got[k], nan=-7.0), torch.nan_to_num(ref[k][:n], nan=-7.0)), (F.__name__, n, k)
            for m in range(3):
                assert torch.equal(got["N_act"][m], ref["N_act"][m][:n]), (F.__name__, n, m)
            assert int(got["n_domain_errors"].item()) == int(torch.isnan(ref["J_hom"][:n]).sum().item()), (F.__name__, n)
        if F is np.float64:
            blk = CMP.pack_icenuc(tps, aps=aip, ap=ap, ad=ad, dust=dust, koop=koop, hom_linear=False)
            o = orc.arg_icenuc(blk, *[st[k][:1000] for k in KEYS])
            got = AA.activation_and_ice_nucleation(ap, ad, aip, tps, dust, koop, *dev(0, 1000), hom_linear=False)
            assert int(got["n_domain_errors"].item()) == o["n_domain_errors"] > 0
        holders = [torch.from_numpy(st[k]).to(cuda) for k in KEYS]
        off = [h[1:] for h in holders]
        assert all(c.data_ptr() % 16 != 0 for c in off)
        ref1 = AA.activation_and_ice_nucleation(ap, ad, aip, tps, dust, koop, *dev(1, n_big + 1), hom_linear=False)
        got1 = AA.activation_and_ice_nucleation(ap, ad, aip, tps, dust, koop, *off, hom_linear=False)
        for k in names:
            assert torch.equal(torch.nan_to_num(got1[k], nan=-7.0), torch.nan_to_num(ref1[k], nan=-7.0)), (F.__name__, "misaligned", k)
        assert int(got1["n_domain_errors"].item()) == int(ref1["n_domain_errors"].item())
