"""GPU parity tests of the 1-moment path (BMT:505-632, CM1, NEQ) through the C-ABI vs the CPU
oracle.  Same criteria as test_gpu_2m.py: Float64 <= 1e-12 relative per output, or (where the
reference algorithm itself cancels) within 2x the reference's own first-order rounding-error
bound; exact zeros (gated regimes) must coincide bit for bit."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
KEYS = ("rho", "T", "q_tot", "q_lcl", "q_icl", "q_rai", "q_sno")
HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "m1_goldens.json")))


def _dev(st, cuda):
    import torch
    return [torch.from_numpy(st[k]).to(cuda) for k in KEYS]


def _call(built, mode, mp, tps, cols, **kw):
    BMT = built.BMT
    return BMT.bulk_microphysics_tendencies(mode, BMT.Microphysics1Moment(), mp, tps, *cols, **kw)


OPTION_SETS = [
    {},
    dict(cloud_ice_formation="TemperatureDependent", rain_autoconversion="PrescribedNd", snow_autoconversion="WithSupersaturation",
         snow_deposition_sublimation="SublimationOnly"),
    dict(rain_snow_accretion=None, snow_melt=None, cloud_ice_melt=None, cloud_liquid_formation=None, cloud_ice_rain_accretion=None),
]


def _opts(CMP, d):
    return {k: (None if v is None else getattr(CMP, v)()) for k, v in d.items()}


@pytest.mark.parametrize("optset", range(len(OPTION_SETS)))
def test_bmt1m_verbose_f64_parity(built, orc, cuda, optset):
    from cumicro.testing import synthetic_states_1m, assert_parity
    CMP, BMT = built.CMP, built.BMT
    n = 1 << 17
    st = synthetic_states_1m(n, seed=1234 + optset)
    mp = CMP.Microphysics1MParams(np.float64, **_opts(CMP, OPTION_SETS[optset]))
    tps = CMP.ThermodynamicsParameters(np.float64)
    blk = CMP.pack_1m(mp, tps)
    out = _call(built, BMT.InstantaneousVerbose(), mp, tps, _dev(st, cuda))
    ref = orc.bmt1m(blk, *[st[k] for k in KEYS], mode="verbose")
    bnd = orc.bmt1m(blk, *[st[k] for k in KEYS], mode="verbose", bound=True)
    for k in orc.OUT_1M + orc.SRC_1M:
        rep = assert_parity(k, out[k].cpu().numpy(), ref[k], bound=bnd[k])
        assert rep["max_rel"] <= 1e-12, (k, rep)
    inst = _call(built, BMT.Instantaneous(), mp, tps, _dev(st, cuda))
    for k in orc.OUT_1M:
        assert np.array_equal(inst[k].cpu().numpy(), out[k].cpu().numpy())   # same kernel maths, fewer stores


@pytest.mark.parametrize("dt,nsub", [(1.0, 1), (60.0, 1), (300.0, 3)])
def test_bmt1m_linearized_average_f64_parity(built, orc, cuda, dt, nsub):
    from cumicro.testing import synthetic_states_1m, assert_parity
    CMP, BMT = built.CMP, built.BMT
    n = 1 << 16
    st = synthetic_states_1m(n, seed=77)
    mp = CMP.Microphysics1MParams(np.float64)
    tps = CMP.ThermodynamicsParameters(np.float64)
    blk = CMP.pack_1m(mp, tps)
    out = _call(built, BMT.LinearizedAverage(), mp, tps, _dev(st, cuda), Δt=dt, nsub=nsub)
    ref = orc.bmt1m(blk, *[st[k] for k in KEYS], mode="linearized_average", dt=dt, nsub=nsub)
    bnd = orc.bmt1m(blk, *[st[k] for k in KEYS], mode="linearized_average", dt=dt, nsub=nsub, bound=True)
    for k in orc.OUT_1M:
        rep = assert_parity(k, out[k].cpu().numpy(), ref[k], bound=bnd[k])
        assert rep["max_rel"] <= 1e-12, (k, rep)


def test_goldens_through_the_gpu(built, cuda):
    import torch
    CMP, BMT = built.CMP, built.BMT
    mp = CMP.Microphysics1MParams(np.float64)
    tps = CMP.ThermodynamicsParameters(np.float64)
    s = G["accretion"]["state"]
    f = lambda v: torch.full((4,), v, dtype=torch.float64, device=cuda)
    q = f(s["q"])
    cold = _call(built, BMT.InstantaneousVerbose(), mp, tps, [f(s["rho"]), f(263.0), f(15e-3), q, q, q, q])
    warm = _call(built, BMT.InstantaneousVerbose(), mp, tps, [f(s["rho"]), f(283.0), f(15e-3), q, q, q, q])
    for k, (val, where) in G["accretion"].items():
        if k == "state":
            continue
        got = float((warm if k.endswith("_warm") else cold)[k][0])
        assert abs(got / val - 1) < 1e-13, (k, got, val, where)
    g = G["snow_melt"]
    z = f(0.0)
    o = _call(built, BMT.InstantaneousVerbose(), mp, tps, [f(g["rho"]), f(273.15 + g["dT"]), f(1e-2), z, z, z, f(g["q_sno"])])
    assert abs(float(o["S_melt_sno_rai"][0]) / g["value"] - 1) < 1e-13
    vels = {"rain_chen": ("rain", CMP.Chen2022VelTypeRain), "snow_chen": ("snow", CMP.Chen2022VelTypeLargeIce),
            "cloud_liquid_stokes": ("cloud_liquid", CMP.StokesRegimeVelType), "cloud_ice_chen": ("cloud_ice", CMP.Chen2022VelTypeSmallIce)}
    for kind, rho, qq, val, where in G["velocities"]:
        sp, V = vels[kind]
        got = float(built.CM1.terminal_velocity(mp, tps, sp, V(np.float64), f(rho), f(qq))[0])
        assert abs(got / val - 1) < 1e-12, (kind, got, val, where)


def test_terminal_velocities_f64_parity(built, orc, cuda):
    import torch
    from cumicro.testing import synthetic_states_1m, assert_parity
    CMP = built.CMP
    st = synthetic_states_1m(1 << 15, seed=5)
    mp = CMP.Microphysics1MParams(np.float64)
    tps = CMP.ThermodynamicsParameters(np.float64)
    blk = CMP.pack_1m(mp, tps)
    rho = torch.from_numpy(st["rho"]).to(cuda)
    cases = [("rain", None, "rain_blk1m", "q_rai"), ("snow", None, "snow_blk1m", "q_sno"),
             ("rain", CMP.Chen2022VelTypeRain(np.float64), "rain_chen", "q_rai"),
             ("snow", CMP.Chen2022VelTypeLargeIce(np.float64), "snow_chen", "q_sno"),
             ("cloud_liquid", CMP.StokesRegimeVelType(np.float64), "cloud_liquid_stokes", "q_lcl"),
             ("cloud_ice", CMP.Chen2022VelTypeSmallIce(np.float64), "cloud_ice_chen", "q_icl")]
    for sp, vel, kind, qk in cases:
        got = built.CM1.terminal_velocity(mp, tps, sp, vel, rho, torch.from_numpy(st[qk]).to(cuda)).cpu().numpy()
        ref = orc.termvel_1m(blk, kind, st["rho"], st[qk], vel)
        bnd = orc.termvel_1m(blk, kind, st["rho"], st[qk], vel, bound=True)
        rep = assert_parity(kind, got, ref, bound=bnd)
        assert rep["max_rel"] <= 1e-12, (kind, rep)


def test_edge_cases(built, orc, cuda):
    """Zeros / negatives in every slot, T exactly at T_freeze (>=, <=, > predicates), n = 0, 1, 3."""
    CMP, BMT = built.CMP, built.BMT
    mp = CMP.Microphysics1MParams(np.float64)
    tps = CMP.ThermodynamicsParameters(np.float64)
    blk = CMP.pack_1m(mp, tps)
    base = dict(rho=1.0, T=270.0, q_tot=8e-3, q_lcl=1e-3, q_icl=5e-4, q_rai=2e-4, q_sno=3e-4)
    rows = [dict(base), dict(base, T=273.15), dict(base, T=np.nextafter(273.15, 300)), dict(base, T=np.nextafter(273.15, 0)), dict(base, T=290.0)]
    for k in ("q_tot", "q_lcl", "q_icl", "q_rai", "q_sno"):
        for v in (0.0, -1e-6):
            rows.append(dict(base, **{k: v}))
            rows.append(dict(base, T=280.0, **{k: v}))
    rows.append(dict(base, q_lcl=0.0, q_icl=0.0, q_rai=0.0, q_sno=0.0))
    st = {k: np.array([r[k] for r in rows]) for k in KEYS}
    out = _call(built, BMT.InstantaneousVerbose(), mp, tps, _dev(st, cuda))
    ref = orc.bmt1m(blk, *[st[k] for k in KEYS], mode="verbose")
    for k in orc.OUT_1M + orc.SRC_1M:
        g = out[k].cpu().numpy()
        assert np.all(np.isfinite(g)), k
        np.testing.assert_allclose(g, ref[k], rtol=1e-11, atol=0, err_msg=k)
        assert np.array_equal(g == 0, ref[k] == 0), k
    for n in (0, 1, 3):
        sub = {k: v[:n].copy() for k, v in st.items()}
        o = _call(built, BMT.Instantaneous(), mp, tps, _dev(sub, cuda))
        assert o["dq_rai_dt"].shape[0] == n
        if n:
            assert np.array_equal(o["dq_rai_dt"].cpu().numpy(), out["dq_rai_dt"].cpu().numpy()[:n])


def test_bmt1m_f32(built, orc, cuda):
    """Float32 method: <= 4 Float32 ULPs from the true value; regime selection = the Float32
    reference's (Float32 thresholds)."""
    from cumicro.testing import synthetic_states_1m, assert_f32_method
    CMP, BMT = built.CMP, built.BMT
    n = 1 << 15
    st32 = synthetic_states_1m(n, seed=21, dtype=np.float32)
    mp32, tps32 = CMP.Microphysics1MParams(np.float32), CMP.ThermodynamicsParameters(np.float32)
    blk32 = CMP.pack_1m(mp32, tps32)
    blk64 = CMP.widen(blk32)
    out = _call(built, BMT.InstantaneousVerbose(), mp32, tps32, _dev(st32, cuda))
    ref32 = orc.bmt1m(blk32, *[st32[k] for k in KEYS], mode="verbose")
    st64 = [st32[k].astype(np.float64) for k in KEYS]
    with orc.f32_thresholds():
        truth = orc.bmt1m(blk64, *st64, mode="verbose")
        bound = orc.bmt1m(blk64, *st64, mode="verbose", bound=True)
    for k in orc.OUT_1M + orc.SRC_1M:
        assert_f32_method("1m:" + k, out[k].cpu().numpy(), ref32[k], truth[k], bound[k], ref_is_f32_oracle=True)


def test_config1_grid_64cubed(built, orc, cuda):
    """BASELINE config 1: the 64x64x64 column grid (262 144 points, flat SoA), full parity."""
    from cumicro.testing import synthetic_states_1m, assert_parity
    CMP, BMT = built.CMP, built.BMT
    n = 64 * 64 * 64
    st = synthetic_states_1m(n, seed=1234)
    mp = CMP.Microphysics1MParams(np.float64)
    tps = CMP.ThermodynamicsParameters(np.float64)
    blk = CMP.pack_1m(mp, tps)
    out = _call(built, BMT.Instantaneous(), mp, tps, _dev(st, cuda))
    ref = orc.bmt1m(blk, *[st[k] for k in KEYS])
    bnd = orc.bmt1m(blk, *[st[k] for k in KEYS], bound=True)
    for k in orc.OUT_1M:
        rep = assert_parity(k, out[k].cpu().numpy(), ref[k], bound=bnd[k])
        assert rep["max_rel"] <= 1e-12 and rep["frac_forward_ok"] > 0.98, (k, rep)


@pytest.mark.parametrize("FT", [np.float64, np.float32])
def test_0m_bit_exact(built, orc, cuda, FT):
    """BMT:658-680: three IEEE operations per point in the method's own float type -> bit-exact against the oracle."""
    import torch
    CMP, BMT = built.CMP, built.BMT
    mp = CMP.Microphysics0MParams(FT)
    tps = CMP.ThermodynamicsParameters(FT)
    rng = np.random.default_rng(4)
    n = 100003
    ql = (rng.random(n) * 4e-3 - 5e-4).astype(FT)
    qi = (rng.random(n) * 4e-3 - 5e-4).astype(FT)
    qvs = (rng.random(n) * 2e-2).astype(FT)
    T = np.full(n, 280.0, FT)
    d = lambda a: torch.from_numpy(a).to(cuda)
    got = BMT.bulk_microphysics_tendencies(BMT.Microphysics0Moment(), mp, tps, d(T), d(ql), d(qi))
    assert np.array_equal(got.dq_tot_dt.cpu().numpy(), orc.bmt0m(mp.precip, ql, qi))
    got = BMT.bulk_microphysics_tendencies(BMT.Microphysics0Moment(), mp, tps, d(T), d(ql), d(qi), d(qvs))
    ref = orc.bmt0m(mp.precip, ql, qi, qvs)
    assert np.array_equal(got.dq_tot_dt.cpu().numpy(), ref) and (ref < 0).mean() > 0.3 and (ref == 0).mean() > 0.01


def test_noneq_leaf_methods(built, orc, cuda):
    """CMNonEq.conv_q_vap_to_q_lcl / conv_q_vap_to_q_icl as stand-alone array methods (NEQ:110-224): the reference's
    literals (test/microphysics_noneq_tests.jl:84-88, rho = 0.8, T = 263, q_tot = 1.2 q_sat) and seeded columns vs the oracle."""
    import json
    import os
    import torch
    from cumicro.testing import psat_liq, psat_ice, synthetic_states_1m, assert_parity
    CMP, NEQ = built.CMP, built.CMNonEq
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "m1_goldens.json")))["noneq"]
    mp, tps = CMP.Microphysics1MParams(np.float64), CMP.ThermodynamicsParameters(np.float64)
    rho, T = g["rho"], g["T"]
    Rv = CMP.DEFAULTS["gas_constant_vapor"]
    col = lambda v: torch.full((16,), v, dtype=torch.float64, device=cuda)
    for fn, opt, psat, want in ((NEQ.conv_q_vap_to_q_lcl, CMP.CloudLiquidFormation(), psat_liq, g["cond"]),
                                (NEQ.conv_q_vap_to_q_icl, CMP.ConstantTimescale(), psat_ice, g["dep"])):
        micro = dict(q_tot=col(1.2 * psat(T) / (Rv * rho * T)), q_lcl=col(0.0), q_icl=col(0.0), q_rai=col(0.0), q_sno=col(0.0))
        out = fn(opt, mp, tps, micro, dict(ρ=col(rho), T=col(T))).cpu().numpy()
        assert np.all(out == out[0]) and abs(out[0] / want - 1) < 1e-12, (out[0], want)
        assert torch.all(fn(None, mp, tps, micro, dict(rho=col(rho), T=col(T))) == 0)
    # seeded columns, both cloud-ice options, vs the oracle's source terms
    n = 1 << 16
    st = synthetic_states_1m(n, seed=99)
    cols = {k: torch.from_numpy(v).to(cuda) for k, v in st.items()}
    micro = {k: cols[k] for k in ("q_tot", "q_lcl", "q_icl", "q_rai", "q_sno")}
    thermo = dict(ρ=cols["rho"], T=cols["T"])
    for opt in (CMP.ConstantTimescale(), CMP.TemperatureDependent()):
        mpo = CMP.Microphysics1MParams(np.float64, cloud_ice_formation=opt)
        ref = orc.bmt1m(CMP.pack_1m(mpo, tps), *[st[k] for k in ("rho", "T", "q_tot", "q_lcl", "q_icl", "q_rai", "q_sno")], mode="verbose")
        bound = orc.bmt1m(CMP.pack_1m(mpo, tps), *[st[k] for k in ("rho", "T", "q_tot", "q_lcl", "q_icl", "q_rai", "q_sno")], mode="verbose", bound=True)
        got = NEQ.conv_q_vap_to_q_icl(opt, mp, tps, micro, thermo).cpu().numpy()   # mp has the default option: the method dispatches on `opt`
        assert_parity("conv_q_vap_to_q_icl", got, ref["S_phase_change_vap_icl"], bound=bound["S_phase_change_vap_icl"])
    got = NEQ.conv_q_vap_to_q_lcl(CMP.CloudLiquidFormation(), mp, tps, micro, thermo).cpu().numpy()
    assert_parity("conv_q_vap_to_q_lcl", got, ref["S_phase_change_vap_lcl"], bound=bound["S_phase_change_vap_lcl"])
    with pytest.raises(TypeError):
        NEQ.conv_q_vap_to_q_icl(CMP.CloudLiquidFormation(), mp, tps, micro, thermo)


def test_tile_shape_ragged_sizes_and_misaligned_columns(built, cuda):
    """The tile launch shape (cm_launch.cuh, pointwise_kernel_tiled): full tiles arrive by bulk copies, the partial last tile and
    columns that are not 16-byte aligned by guarded scalar loads — a point's bits do not depend on which way it was loaded, on
    its position in a tile, or on the size of the call."""
    import torch
    from cumicro.testing import synthetic_states_1m
    CMP, BMT = built.CMP, built.BMT
    mp, tps = CMP.Microphysics1MParams(np.float64), CMP.ThermodynamicsParameters(np.float64)
    n_big = 128 * 37 + 5
    st = synthetic_states_1m(n_big + 1, seed=77)
    full = _dev({k: v[:n_big].copy() for k, v in st.items()}, cuda)
    for mode in (BMT.Instantaneous(), BMT.InstantaneousVerbose()):
        ref = _call(built, mode, mp, tps, full)
        keys = list(ref.keys()) if hasattr(ref, "keys") else ["dq_lcl_dt", "dq_icl_dt", "dq_rai_dt", "dq_sno_dt"]
        for n in (1, 127, 128, 129, 1000, n_big):
            sub = _dev({k: v[:n].copy() for k, v in st.items()}, cuda)
            got = _call(built, mode, mp, tps, sub)
            for k in keys:
                assert torch.equal(got[k], ref[k][:n]), (type(mode).__name__, n, k)
        # one element into a 16-byte aligned allocation: 8-byte aligned only
        holders = _dev(st, cuda)
        off = [h[1:] for h in holders]
        assert all(c.data_ptr() % 16 == 8 for c in off)
        ref1 = _call(built, mode, mp, tps, _dev({k: v[1:].copy() for k, v in st.items()}, cuda))
        got1 = _call(built, mode, mp, tps, off)
        for k in keys:
            assert torch.equal(got1[k], ref1[k]), (type(mode).__name__, "misaligned", k)


def test_non_default_exponent_structure_runs_the_generic_body(built, orc, cuda):
    """The default 1-moment exponents (quarter / eighth powers of λ⁻¹) run the body that forms every power by multiplication
    (cm_1m.cuh, STD); any other exponent set runs the generic exp-based body.  Both against the oracle, on the same states, and a
    block whose only non-default entry is a Δ of 0 must reproduce the default bits (it IS the default structure)."""
    from cumicro.testing import synthetic_states_1m, assert_parity
    CMP, BMT = built.CMP, built.BMT
    tps = CMP.ThermodynamicsParameters(np.float64)
    n = 1 << 15
    st = synthetic_states_1m(n, seed=4321)
    cols = _dev(st, cuda)
    sets = [{"rain_cross_section_size_relation_coefficient_dela": 0.07, "snow_terminal_velocity_size_relation_coefficient_delv": -0.03},
            {"rain_mass_size_relation_coefficient_delm": 0.11, "snow_mass_size_relation_coefficient_delm": -0.2,
             "rain_terminal_velocity_size_relation_coefficient_delv": 0.05}]
    for ov in sets:
        mp = CMP.Microphysics1MParams(np.float64, overrides=ov)
        blk = CMP.pack_1m(mp, tps)
        out = _call(built, BMT.InstantaneousVerbose(), mp, tps, cols)
        ref = orc.bmt1m(blk, *[st[k] for k in KEYS], mode="verbose")
        bnd = orc.bmt1m(blk, *[st[k] for k in KEYS], mode="verbose", bound=True)
        for k in orc.OUT_1M + orc.SRC_1M:
            rep = assert_parity(k, out[k].cpu().numpy(), ref[k], bound=bnd[k])
            assert rep["max_rel"] <= 1e-12, (ov, k, rep)
        la = _call(built, BMT.LinearizedAverage(), mp, tps, cols, Δt=60.0, nsub=2)
        rl = orc.bmt1m(blk, *[st[k] for k in KEYS], mode="linearized_average", dt=60.0, nsub=2)
        bl = orc.bmt1m(blk, *[st[k] for k in KEYS], mode="linearized_average", dt=60.0, nsub=2, bound=True)
        for k in orc.OUT_1M:
            rep = assert_parity("linavg " + k, la[k].cpu().numpy(), rl[k], bound=bl[k])
            assert rep["max_rel"] <= 1e-12, (ov, k, rep)
    base = _call(built, BMT.Instantaneous(), CMP.Microphysics1MParams(np.float64), tps, cols)
    same = _call(built, BMT.Instantaneous(), CMP.Microphysics1MParams(np.float64, overrides={"rain_mass_size_relation_coefficient_delm": 0.0}), tps, cols)
    for k in orc.OUT_1M:
        assert np.array_equal(base[k].cpu().numpy(), same[k].cpu().numpy())
