"""GPU parity of the P3 kernels through the C-ABI vs the CPU oracle (BASELINE config 4 and the
2-moment + P3 fused tendencies, BMT:898-1083).  Float64: 1e-12 relative, or the reference
algorithm's own first-order rounding bound where it subtracts nearly equal numbers
(cumicro.testing.compare_report); regime selection (exact zeros, non-finite values) bit-exact."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "p3_goldens.json")))
EPS = np.finfo(np.float64).eps
IN12 = ("rho", "T", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice", "q_rim", "b_rim")
RATE2ORC = dict(v_n="v_n", v_m="v_m", melt_dNdt="melt_dN", melt_dLdt="melt_dL", self_collection_dNdt="selfcol", dq_c="dq_c", dq_r="dq_r",
                dN_c="dN_c", dN_r="dN_r", dL_rim="dL_rim", dL_ice="dL_ice", dB_rim="dB_rim")


def _setup(built, n, seed=1234, **mpkw):
    CMP, T_ = built.CMP, built.testing
    mp = CMP.Microphysics2MParams(np.float64, with_ice=True, **mpkw)
    tps = CMP.ThermodynamicsParameters(np.float64)
    st = T_.synthetic_states_p3(n, seed=seed)
    return mp, tps, st


def _volumetric(st):
    return [st[k] * st["rho"] for k in ("q_ice", "n_ice", "q_rim", "b_rim")]


def _converged_logl(orc, blk, st):
    l = orc.p3_state(blk, *_volumetric(st), from_prognostic=True, want=("logl",), logl_iters=40)["logl"]
    return np.where(np.isfinite(l), l, 0.0)


def _oracle_rates(orc, blk, st, logl, sel, bound=False):
    rho = st["rho"]
    o = orc.p3_state(blk, *[v[sel] for v in _volumetric(st)], from_prognostic=True, rho_a=rho[sel], T=st["T"][sel], logl=logl[sel],
                     L_c=(st["q_lcl"] * rho)[sel], N_c=(st["n_lcl"] * rho)[sel], L_r=(st["q_rai"] * rho)[sel], N_r=(st["n_rai"] * rho)[sel],
                     want=("v_n", "v_m", "melt", "selfcol", "src7"), bound=bound)
    warm = st["T"][sel] > 273.15     # BMT:981-985 evaluates ice_melt only above freezing (below, max(0, .) gives 0 anyway)
    for k in ("melt_dN", "melt_dL"):
        o[k] = np.where(warm, o[k], 0.0)
    return o


@pytest.mark.parametrize("variant", ["default_gl16", "cheb20_unlimited", "gl12_noar_constslope"])
def test_p3_rates_f64_parity(built, orc, cuda, variant):
    import torch
    from cumicro.testing import assert_parity
    P3, CMP3 = built.P3, built.CMP3
    kw, quad, ice_kw = {}, None, {}
    if variant == "cheb20_unlimited":
        kw = dict(is_limited=False, quadrature_order=20)          # build_quadrature(20) -> ChebyshevGauss
    # 2^14 points for the default configuration (BASELINE config 4's scheme), 2^11 for the two variants (the CPU port does ~1e4 points/s)
    mp, tps, st = _setup(built, (1 << 14) if variant == "default_gl16" else (1 << 11), seed=11 + len(variant), **kw)
    if variant == "gl12_noar_constslope":
        mp.ice = CMP3.P3IceParams(np.float64, slope_law="constant", aspect_ratio=CMP3.NoAspectRatio(), quadrature_order=12,
                                  overrides={"P3_constant_slope_parameterization_value": 1.5})
        quad = CMP3.GaussLegendre(np.float64, 12)
    blk = CMP3.pack_p3(mp, tps, quad=quad)
    logl = _converged_logl(orc, blk, st)
    d = {k: torch.from_numpy(v).to(cuda) for k, v in st.items()}
    got = P3.process_rates(mp, tps, *[d[k] for k in IN12], torch.from_numpy(logl).to(cuda), quad=quad)
    ice = (st["q_ice"] > EPS) & (st["n_ice"] > EPS)
    assert 0.5 < ice.mean() < 0.9
    ref, bnd = _oracle_rates(orc, blk, st, logl, ice), _oracle_rates(orc, blk, st, logl, ice, bound=True)
    F_rim0 = (st["q_rim"][ice] == 0).mean()
    assert F_rim0 > 0.05                                            # the unrimed regime (Inf thresholds) is populated
    worst = 0.0
    for g, r in RATE2ORC.items():
        gg = got[g].cpu().numpy()
        rep = assert_parity(f"{variant}:{g}", gg[ice], ref[r], bound=bnd[r])
        worst = max(worst, rep["max_rel"])
        if g not in ("v_n", "v_m"):
            assert np.all(gg[~ice] == 0), g                        # BMT:961 branch not taken
    vel_off = (st["n_ice"] * st["rho"] < EPS) | (st["q_ice"] * st["rho"] < EPS)
    assert np.all(got["v_n"].cpu().numpy()[vel_off] == 0) and np.all(got["v_m"].cpu().numpy()[vel_off] == 0)
    assert worst < 1e-12


def test_bmt2m_p3_f64_parity(built, orc, cuda):
    import torch
    from cumicro.testing import assert_parity
    BMT, CMP3 = built.BMT, built.CMP3
    mp, tps, st = _setup(built, 1 << 13, seed=5)
    # cold points so that F23 deposition / Bigg freezing / immersion cap are all active somewhere
    st["T"][::3] -= 25.0
    blk = CMP3.pack_p3(mp, tps)
    logl = _converged_logl(orc, blk, st)
    shift = np.random.default_rng(1).normal(0, 1.0, st["rho"].size)
    cols = [st[k] for k in orc.P3_BMT_IN[:-1]] + [logl]
    ref = orc.bmt2m_p3(blk, *cols, inpc_log_shift=shift)
    bnd = orc.bmt2m_p3(blk, *cols, inpc_log_shift=shift, bound=True)
    dcols = [torch.from_numpy(c).to(cuda) for c in cols]
    got = BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp, tps, *dcols, torch.from_numpy(shift).to(cuda))
    for k in orc.P3_BMT_OUT[:-1]:
        assert_parity(k, got[k].cpu().numpy(), ref[k], bound=bnd[k])
    assert float(got["dn_lcl_activation_dt"].abs().max()) == 0.0
    # regimes exercised: ice-free points, melting points, freezing points, deposition nucleation
    ice = (st["q_ice"] > EPS) & (st["n_ice"] > EPS)
    assert (~ice).sum() > 100 and (ice & (st["T"] > 273.15)).sum() > 20 and (st["T"] < 258.15).sum() > 100
    assert (ref["dq_rim_dt"] != 0).mean() > 0.3
    # without the optional shift column (inpc_log_shift = 0, BMT:903)
    ref0 = orc.bmt2m_p3(blk, *cols)
    got0 = BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp, tps, *dcols)
    bnd0 = orc.bmt2m_p3(blk, *cols, bound=True)
    for k in ("dq_ice_dt", "dn_ice_dt", "dq_lcl_dt"):
        assert_parity(k + ":noshift", got0[k].cpu().numpy(), ref0[k], bound=bnd0[k])


def test_p3_goldens_through_the_gpu(built, cuda):
    """The reference's own P3 literals (tests/golden/p3_goldens.json) through the C-ABI."""
    import torch
    P3, CMP, CMP3 = built.P3, built.CMP, built.CMP3
    mp = CMP.Microphysics2MParams(np.float64, with_ice=True)
    tps = CMP.ThermodynamicsParameters(np.float64)
    q12 = CMP3.GaussLegendre(np.float64, 12)
    f = lambda v: torch.full((3,), float(v), dtype=torch.float64, device=cuda)
    g = G["bulk_velocity"]
    for k, F in enumerate(g["F_rims"]):
        L, N = g["L_ice"], g["N_ice"]
        L_rim = F * L
        B_rim = L_rim / g["rho_rim"]
        logl = P3.get_distribution_logλ_from_prognostic(mp, tps, f(L), f(N), f(L_rim), f(B_rim))
        v_n, v_m = P3.ice_terminal_velocities_from_prognostic(mp, tps, f(g["rho_a"]), f(L), f(N), f(L_rim), f(B_rim), logl, quad=q12)
        # state_from_prognostic regularises F_rim = L_rim / L: equal to the literal F_rim to rounding
        assert abs(float(v_n[0]) / g["v_n_phi"][k] - 1) < 1e-12
        assert abs(float(v_m[0]) / g["v_m_phi"][k] - 1) < 1e-12
    s = G["process_state"]
    rho = s["rho_a"]
    q_rim = s["F_rim"] * s["q_ice"]
    b_rim = q_rim / s["rho_rim"]
    Tf = 273.15
    c = G["collisions"]
    logl = P3.get_distribution_logλ_from_prognostic(mp, tps, f(s["q_ice"] * rho), f(s["n_ice"] * rho), f(q_rim * rho), f(b_rim * rho))
    for m in G["melt"]:
        r = P3.process_rates(mp, tps, f(rho), f(Tf + m["dT"]), f(0), f(0), f(0), f(0), f(s["q_ice"]), f(s["n_ice"]), f(q_rim), f(b_rim), logl,
                             quad=q12, which=("melt_dNdt", "melt_dLdt"))
        assert abs(float(r.melt_dNdt[0]) / m["dNdt"] - 1) < 1e-11
        assert abs(float(r.melt_dLdt[0]) / m["dLdt"] - 1) < 1e-11
    r = P3.process_rates(mp, tps, f(rho), f(Tf + c["dT"]), f(c["L_c"] / rho), f(c["N_c"] / rho), f(c["L_r"] / rho), f(c["N_r"] / rho),
                         f(s["q_ice"]), f(s["n_ice"]), f(q_rim), f(b_rim), logl, quad=q12)
    v = c["values"]
    # ∂ₜL_ice = QCFRZ + QRFRZ, ∂ₜq_c = -(QCFRZ + QCSHD)/ρ ... (P3_processes.jl:640-650); the cloud literals are stale at 4.5e-4
    assert abs(float(r.dL_ice[0]) / (v["QCFRZ"] + v["QRFRZ"]) - 1) < 1e-5
    assert abs(float(r.dq_c[0]) / (-(v["QCFRZ"] + v["QCSHD"]) / rho) - 1) < c["rtol"]
    assert abs(float(r.dN_c[0]) / (-v["NCCOL"]) - 1) < c["rtol"]
    assert float(r.self_collection_dNdt[0]) > 0


def test_p3_logl_solver_parity(built, orc, cuda):
    import torch
    P3, CMP3 = built.P3, built.CMP3
    mp, tps, st = _setup(built, 4000, seed=3)
    blk = CMP3.pack_p3(mp, tps)
    vol = _volumetric(st)
    ref = orc.p3_state(blk, *vol, from_prognostic=True, want=("logl",))["logl"]
    dv = [torch.from_numpy(v).to(cuda) for v in vol]
    got = P3.get_distribution_logλ_from_prognostic(mp, tps, *dv).cpu().numpy()
    fin = np.isfinite(ref)
    assert np.array_equal(np.isneginf(got), np.isneginf(ref)) and (~fin).sum() > 500     # empty ice -> log(0)  (:289)
    # same fixed 10 Brent iterations as the oracle: the iterates agree to rounding except where a branch of
    # Brent's method flips on a rounding-level tie (both then sit within the solver's own residual error)
    d = np.abs(got[fin] - ref[fin])
    assert np.mean(d < 1e-10) > 0.99, np.mean(d < 1e-10)
    conv = orc.p3_state(blk, *vol, from_prognostic=True, want=("logl",), logl_iters=40)["logl"]
    assert np.max(np.abs(got[fin] - conv[fin])) <= np.max(np.abs(ref[fin] - conv[fin])) + 1e-6
    got40 = P3.get_distribution_logλ_from_prognostic(mp, tps, *dv, brent_iters=40).cpu().numpy()
    assert np.max(np.abs(got40[fin] - conv[fin])) < 1e-9


@pytest.mark.parametrize("n", [0, 1, 31, 33, 100])
def test_p3_ragged_sizes_and_slab_independence(built, orc, cuda, n):
    import torch
    P3, CMP3, BMT = built.P3, built.CMP3, built.BMT
    mp, tps, st = _setup(built, 100, seed=9)
    blk = CMP3.pack_p3(mp, tps)
    logl = _converged_logl(orc, blk, st)
    cols = [torch.from_numpy(st[k]).to(cuda) for k in orc.P3_BMT_IN[:-1]] + [torch.from_numpy(logl).to(cuda)]
    full = BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp, tps, *cols)
    part = BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp, tps, *[c[:n].contiguous() for c in cols])
    for k in orc.P3_BMT_OUT[:-1]:
        assert part[k].shape[0] == n
        assert torch.equal(part[k], full[k][:n]), k            # a point's result does not depend on its tile mates
    if n:   # an offset slab (different tile alignment) gives the same bits
        off = BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp, tps, *[c[100 - n:].contiguous() for c in cols])
        for k in orc.P3_BMT_OUT[:-1]:
            assert torch.equal(off[k], full[k][100 - n:]), k


def test_p3_api_errors(built, cuda):
    import torch
    CMP, BMT, CMP3 = built.CMP, built.BMT, built.CMP3
    mp = CMP.Microphysics2MParams(np.float64, with_ice=True)
    tps = CMP.ThermodynamicsParameters(np.float64)
    x = torch.ones(8, dtype=torch.float64, device=cuda)
    with pytest.raises(built._abi.CuMicroError):
        BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp, tps, *[torch.ones(8, dtype=torch.float64)] * 12)   # CPU tensors
    bad = CMP3.GaussLegendre(np.float64, 12)
    bad.n = 0
    with pytest.raises(built._abi.CuMicroError):
        BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp, tps, *[x] * 12, quad=bad)
    with pytest.raises(ValueError):
        CMP3.GaussLegendre(np.float64, 200)


def test_p3_f32_methods(built, orc, cuda):
    """Float32 methods: Float32 columns and parameters, Float64 arithmetic with the Float32
    method's thresholds and iteration counts (eps(Float32) gates, 20 gamma_inc terms, 8 Brent steps), one rounding
    on store.  Criterion: <= 4 Float32 ULP from the true value of the Float32 method (the Float64 oracle in
    f32_thresholds mode on the same Float32 inputs and widened parameters)."""
    import torch
    from cumicro.testing import assert_f32_method
    CMP, CMP3, P3, BMT, T_ = built.CMP, built.CMP3, built.P3, built.BMT, built.testing
    F = np.float32
    mp = CMP.Microphysics2MParams(F, with_ice=True)
    tps = CMP.ThermodynamicsParameters(F)
    blk64 = CMP.widen(CMP3.pack_p3(mp, tps))
    st = T_.synthetic_states_p3(700, seed=21, dtype=F)
    st["T"][::3] -= F(25.0)
    w = {k: v.astype(np.float64) for k, v in st.items()}
    # Float32 arithmetic of BMT:921-929 for the volumetric quantities
    vol32 = [(st[k] * st["rho"]).astype(np.float64) for k in ("q_ice", "n_ice", "q_rim", "b_rim")]
    e32 = np.finfo(F).eps
    with orc.f32_thresholds():
        l = orc.p3_state(blk64, *vol32, from_prognostic=True, want=("logl",), logl_iters=40)["logl"]
    logl = np.where(np.isfinite(l), l, 0.0).astype(F)
    dcols = [torch.from_numpy(st[k]).to(cuda) for k in orc.P3_BMT_IN[:-1]] + [torch.from_numpy(logl).to(cuda)]
    got = BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp, tps, *dcols)
    assert got["dq_ice_dt"].dtype == torch.float32
    # truth: the volumetric products are formed in Float64 by the library from the Float32 specific quantities, as
    # the Float64 oracle does from the widened inputs
    cols64 = [w[k] for k in orc.P3_BMT_IN[:-1]] + [logl.astype(np.float64)]
    with orc.f32_thresholds():
        truth = orc.bmt2m_p3(blk64, *cols64)
        bound = orc.bmt2m_p3(blk64, *cols64, bound=True)
    worst = 0.0
    for k in orc.P3_BMT_OUT[:-1]:
        g = got[k].cpu().numpy()
        worst = max(worst, assert_f32_method(k, g, truth[k].astype(F), truth[k], bound[k]))
    ice = (st["q_ice"] > e32) & (st["n_ice"] > e32)
    assert 0.5 < ice.mean() < 0.9 and worst <= 4
    # logλ solver, Float32 method: 8 Brent iterations
    dv = [torch.from_numpy(v.astype(F)).to(cuda) for v in vol32]
    gl = P3.get_distribution_logλ_from_prognostic(mp, tps, *dv).cpu().numpy()
    with orc.f32_thresholds():
        rl = orc.p3_state(blk64, *[v.astype(F).astype(np.float64) for v in vol32], from_prognostic=True, want=("logl",))["logl"]
    fin = np.isfinite(rl)
    assert np.array_equal(np.isneginf(gl), np.isneginf(rl))
    # 8 Brent iterations are not converged everywhere (tests/test_oracle_p3.py): where a branch of Brent's method flips on
    # a rounding-level tie the two 8th iterates differ, both within the solver's own residual of the converged root
    with orc.f32_thresholds():
        conv = orc.p3_state(blk64, *[v.astype(F).astype(np.float64) for v in vol32], from_prognostic=True, want=("logl",), logl_iters=40)["logl"]
    close = np.abs(gl[fin] - rl[fin]) <= 4 * np.spacing(np.abs(rl[fin]).astype(F))
    assert np.mean(close) > 0.99, np.mean(close)
    assert np.max(np.abs(gl[fin] - conv[fin])) <= np.max(np.abs(rl[fin] - conv[fin])) + 1e-5


def test_p3_full_size_properties_2pow22(built, cuda):
    """BASELINE config 4 size (2^22 points, Float64): size-independent properties of the P3 process rates.
    * liquid + ice mass is conserved by the collisions: ρ (∂ₜq_c + ∂ₜq_r) + ∂ₜL_ice = 0  (P3_processes.jl:640-650)
    * velocities, melt and self-collection are non-negative and finite; outside the BMT:961 gate the rates are exactly 0
    * a slab of the grid evaluated on its own gives the same bits (no coupling between points / tiles)."""
    import torch
    CMP, P3, T_ = built.CMP, built.P3, built.testing
    n = 1 << 22
    mp = CMP.Microphysics2MParams(np.float64, with_ice=True)
    tps = CMP.ThermodynamicsParameters(np.float64)
    st = T_.synthetic_states_p3(n, seed=1234)
    d = {k: torch.from_numpy(v).to(cuda) for k, v in st.items()}
    vol = [d[k] * d["rho"] for k in ("q_ice", "n_ice", "q_rim", "b_rim")]
    logl = P3.get_distribution_logλ_from_prognostic(mp, tps, *vol)
    ice = (d["q_ice"] > EPS) & (d["n_ice"] > EPS)
    assert bool(torch.isfinite(logl[ice]).all()) and bool(torch.isneginf(logl[(vol[0] < EPS) | (vol[1] < EPS)]).all())
    assert float(logl[ice].min()) >= 2.0 and float(logl[ice].max()) <= 17.0
    logl = torch.where(torch.isfinite(logl), logl, torch.zeros_like(logl))
    cols = [d[k] for k in IN12] + [logl]
    r = P3.process_rates(mp, tps, *cols)
    for k, v in r.items():
        assert bool(torch.isfinite(v).all()), k
    for k in ("v_n", "v_m", "melt_dNdt", "melt_dLdt", "self_collection_dNdt", "dL_ice", "dL_rim"):
        assert float(r[k].min()) >= 0.0, k
    for k in ("dq_c", "dN_c"):
        assert float(r[k].max()) <= 0.0, k
    for k in P3.RATE_NAMES[2:]:
        assert float(r[k][~ice].abs().max()) == 0.0, k
    assert float(r["v_m"][ice].min()) > 0.0 and float((r["v_m"] >= r["v_n"])[ice].double().mean()) > 0.99
    resid = d["rho"] * (r["dq_c"] + r["dq_r"]) + r["dL_ice"]
    scale = r["dL_ice"].abs() + (d["rho"] * r["dq_r"]).abs() + (d["rho"] * r["dq_c"]).abs()
    assert float((resid.abs() / torch.clamp(scale, min=1e-300)).max()) < 1e-14
    assert float((r["melt_dLdt"][ice & (d["T"] <= 273.15)]).abs().max()) == 0.0
    lo, hi = n // 3 + 5, n // 3 + 5 + (1 << 16) + 7        # a slab that is not tile-aligned
    part = P3.process_rates(mp, tps, *[c[lo:hi].contiguous() for c in cols])
    for k in P3.RATE_NAMES:
        assert torch.equal(part[k], r[k][lo:hi]), k


def test_f23_and_bigg_rates_standalone(built, orc, cuda):
    """IN.liquid_freezing_rate (rain / cloud PSD), immersion_limit_rate, deposition_rate as BMT:998-1075 calls them."""
    import torch
    from cumicro.testing import assert_parity
    IN, CMP3 = built.IN, built.CMP3
    mp, tps, st = _setup(built, 6000, seed=31)
    st["T"][::2] -= 30.0
    blk = CMP3.pack_p3(mp, tps)
    keys = ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice")
    shift = np.random.default_rng(2).normal(0, 1.0, st["rho"].size)
    d = [torch.from_numpy(st[k]).to(cuda) for k in keys]
    for sh in (None, shift):
        got = IN.f23_and_bigg_rates(mp, tps, *d, None if sh is None else torch.from_numpy(sh).to(cuda))
        ref = orc.icenuc_f23(blk, *[st[k] for k in keys], inpc_log_shift=sh)
        bnd = orc.icenuc_f23(blk, *[st[k] for k in keys], inpc_log_shift=sh, bound=True)
        for k in IN.F23_OUT:
            assert_parity(k, got[k].cpu().numpy(), ref[k], bound=bnd[k])
        assert (ref["rain_dn_frz"] > 0).mean() > 0.2 and (ref["cloud_dq_frz"] > 0).mean() > 0.2 and (ref["deposition_dn"] > 0).mean() > 0.005
        assert (ref["immersion_limit_dn"] == 0).mean() > 0.1           # T >= T_freeze (IN:425)


def test_p3_state_and_shared_numerics(built, orc, cuda):
    """Thresholds / D_m and UT.gamma_inc / gamma_inc_inv / regularised ratios on the device vs the oracle (and scipy,
    at the reference's own tolerances, test/gpu_tests.jl:1305-1338)."""
    import torch
    from cumicro.testing import assert_parity
    sp = pytest.importorskip("scipy.special")
    P3, CMP3 = built.P3, built.CMP3
    mp, tps, st = _setup(built, 3000, seed=41)
    blk = CMP3.pack_p3(mp, tps)
    vol = _volumetric(st)
    ice = (vol[0] > EPS) & (vol[1] > EPS)
    logl = _converged_logl(orc, blk, st)
    d = [torch.from_numpy(v).to(cuda) for v in vol] + [torch.from_numpy(logl).to(cuda)]
    got = P3.state_from_prognostic(mp, tps, *d)
    ref = orc.p3_state(blk, *[v[ice] for v in vol], from_prognostic=True, logl=logl[ice], want=("thresholds", "D_m"))
    for g, r in (("F_rim", "F_rim"), ("ρ_g", "rho_g"), ("D_th", "D_th"), ("D_gr", "D_gr"), ("D_cr", "D_cr"), ("D_m", "D_m")):
        gg = got[g].cpu().numpy()[ice]
        rr = ref[r]
        if g == "ρ_g":   # NaN for unrimed ice in both (P3_particle_properties.jl:33)
            assert np.array_equal(np.isnan(gg), np.isnan(rr)) and np.isnan(rr).sum() > 50
        assert_parity(g, gg, rr)
    assert np.isinf(ref["D_gr"]).sum() > 50
    g = G["bulk_velocity"]
    f = lambda v: torch.full((2,), float(v), dtype=torch.float64, device=cuda)
    for k, F in enumerate(g["F_rims"]):
        L, N = g["L_ice"], g["N_ice"]
        ll = P3.get_distribution_logλ_from_prognostic(mp, tps, f(L), f(N), f(F * L), f(F * L / g["rho_rim"]))
        assert abs(float(P3.D_m(mp, tps, f(L), f(N), f(F * L), f(F * L / g["rho_rim"]), ll)[0]) / g["D_m"][k] - 1) < 1e-12   # p3_tests.jl:440
    # gamma_inc / gamma_inc_inv
    gg = G["gamma_inc_grid"]
    a, x = np.meshgrid(np.array(gg["a"], float), np.array(gg["x"], float), indexing="ij")
    t = lambda v: torch.from_numpy(np.ascontiguousarray(v.ravel())).to(cuda)
    P, Q = P3.gamma_inc(t(a), t(x))
    assert np.max(np.abs(P.cpu().numpy() - sp.gammainc(a.ravel(), x.ravel()))) < gg["atol_PQ"]
    assert np.max(np.abs(Q.cpu().numpy() - sp.gammaincc(a.ravel(), x.ravel()))) < gg["atol_PQ"]
    a, pq = np.meshgrid(np.array(gg["a"], float), np.array(gg["p"], float), indexing="ij")
    xi = P3.gamma_inc_inv(t(a), t(pq)).cpu().numpy()
    assert np.max(np.abs(xi / sp.gammaincinv(a.ravel(), pq.ravel()) - 1)) < gg["rtol_inv"]
    rng = np.random.default_rng(5)
    a, x = rng.uniform(1.0, 12, 20000), rng.uniform(0, 40, 20000)
    assert_parity("gamma_inc P", P3.gamma_inc(t(a), t(x))[0].cpu().numpy(), orc.p3_leaf("gamma_inc_P", a, x), bound=np.full(a.size, 4 * EPS))
    pq = rng.uniform(1e-6, 1 - 1e-6, 20000)
    assert_parity("gamma_inc_inv", P3.gamma_inc_inv(t(a), t(pq)).cpu().numpy(), orc.p3_leaf("gamma_inc_inv", a, pq), rtol=1e-11)
    q_ice = 10 ** rng.uniform(-18, -3, 20000)
    q_rim = q_ice * rng.uniform(0, 1.2, 20000)
    assert_parity("rime_mass_fraction", P3.rime_mass_fraction(t(q_rim), t(q_ice)).cpu().numpy(), orc.p3_leaf("rime_mass_fraction", q_rim, q_ice), rtol=1e-11)
    assert_parity("rime_density", P3.rime_density(t(q_rim), t(q_ice)).cpu().numpy(), orc.p3_leaf("rime_density", q_rim, q_ice), rtol=1e-11)


def test_log_lambda_solved_inside_the_call(built, cuda):
    """§8(f)-1: with the logλ column NULL the P3 kernels solve get_distribution_logλ_from_prognostic themselves (one listed point per
    thread, before the quantile phase).  Float64: bit-identical to passing the stand-alone solve's column; Float32: the in-kernel
    value is not rounded to Float32 on the way, so the results agree to Float32 rounding of logλ's effect."""
    import torch
    from cumicro import BMT, CMP, P3
    from cumicro.testing import synthetic_states_p3
    n = 5000
    st = synthetic_states_p3(n, seed=11)
    mp3, tps = CMP.Microphysics2MParams(np.float64, with_ice=True), CMP.ThermodynamicsParameters(np.float64)
    d = {k: torch.from_numpy(v).to(cuda) for k, v in st.items()}
    vol = [d[k] * d["rho"] for k in ("q_ice", "n_ice", "q_rim", "b_rim")]
    logl = P3.get_distribution_logλ_from_prognostic(mp3, tps, *vol)
    KP = ("rho", "T", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice", "q_rim", "b_rim")
    a = P3.process_rates(mp3, tps, *[d[k] for k in KP], logl)
    b = P3.process_rates(mp3, tps, *[d[k] for k in KP], None)
    for k in a.keys():
        assert torch.equal(torch.nan_to_num(a[k], nan=-1.0), torch.nan_to_num(b[k], nan=-1.0)), k
    va = P3.ice_terminal_velocities_from_prognostic(mp3, tps, d["rho"], *vol, logl)
    vb = P3.ice_terminal_velocities_from_prognostic(mp3, tps, d["rho"], *vol, None)
    assert torch.equal(va[0], vb[0]) and torch.equal(va[1], vb[1])
    cols = [d[k] for k in ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice", "q_rim", "b_rim")]
    ta = BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp3, tps, *cols, logl)
    tb = BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp3, tps, *cols, None)
    for k in ("dq_ice_dt", "dn_ice_dt", "dq_rim_dt", "db_rim_dt", "dq_rai_dt", "dq_lcl_dt"):
        assert torch.equal(torch.nan_to_num(ta[k], nan=-1.0), torch.nan_to_num(tb[k], nan=-1.0)), k
    assert float(a["v_n"].abs().max()) > 0
