"""Synthetic atmospheric states and parity metrics shared by tests/, bench.py and
``__graft_entry__.smoke()``.  Pure numpy; nothing here touches the CUDA library or
the oracle.

States follow the reference's own generator ``generate_atmospheric_states``
(test/gpu_performance.jl:80-136): a 0-15 km column with a decaying temperature
profile, RH swept 0.05..1.05, condensate = saturation excess + U(0,1)*1e-4.  Julia's
MersenneTwister stream is not reproducible here, so PCG64 with the same seed (1234)
is used and both the CPU and GPU paths always see the same generated bits."""
from __future__ import annotations

import numpy as np

from .parameters import DEFAULTS

F64_RTOL = 1e-12   # north-star Float64 tolerance (relative)
F32_ULPS = 4       # north-star Float32 tolerance (ULP-equivalent)


def _psat(T, LH0, dcp):
    d = DEFAULTS
    Rv, T0, Tt, pt = d["gas_constant_vapor"], d["thermodynamics_temperature_reference"], d["temperature_triple_point"], d["pressure_triple_point"]
    return pt * (T / Tt) ** (dcp / Rv) * np.exp((LH0 - dcp * T0) / Rv * (1 / Tt - 1 / T))


def psat_liq(T):
    d = DEFAULTS
    return _psat(T, d["latent_heat_vaporization_at_reference"], d["isobaric_specific_heat_vapor"] - d["isobaric_specific_heat_liquid"])


def psat_ice(T):
    d = DEFAULTS
    return _psat(T, d["latent_heat_sublimation_at_reference"], d["isobaric_specific_heat_vapor"] - d["isobaric_specific_heat_ice"])


def atmospheric_profile(n, rng):
    """(T, p, rho, q_vap, q_sat_liq, q_sat_ice) for n points, test/gpu_performance.jl:80-118."""
    d = DEFAULTS
    Rd, Rv, g = d["gas_constant_dry_air"], d["gas_constant_vapor"], d["gravitational_acceleration"]
    z = np.linspace(0.0, 15000.0, n)
    T_surf, T_min, p0 = 300.0, 215.0, 1.0e5
    lapse = 6.5e-3
    T = np.maximum(T_surf - lapse * z, T_min)
    # hydrostatic pressure of the (dry) profile above
    z_tp = (T_surf - T_min) / lapse
    p = np.where(z <= z_tp, p0 * (T / T_surf) ** (g / (Rd * lapse)),
                 p0 * (T_min / T_surf) ** (g / (Rd * lapse)) * np.exp(-g * (z - z_tp) / (Rd * T_min)))
    RH = np.linspace(0.05, 1.05, n)
    # decorrelate RH from height so every regime occurs at every temperature
    RH = RH[rng.permutation(n)] if n > 1 else RH
    pv = RH * psat_liq(T)
    eps_ = Rd / Rv
    q_vap = eps_ * pv / (p - (1 - eps_) * pv)
    Rm = Rd * (1 + (Rv / Rd - 1) * q_vap)
    rho = p / (Rm * T)
    q_sat_liq = psat_liq(T) / (rho * Rv * T)
    q_sat_ice = psat_ice(T) / (rho * Rv * T)
    return T, p, rho, q_vap, q_sat_liq, q_sat_ice


def synthetic_states_2m(n, seed=1234, dtype=np.float64, number="loguniform", frac_empty=0.05):
    """Inputs of the 2-moment warm-rain method: rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai.

    ``number='const'`` is the reference's benchmark choice (n_lcl=1e8, n_rai=1e4,
    test/gpu_performance.jl:219-220); ``'loguniform'`` (default) draws
    n_lcl in [1e7,1e9], n_rai in [1e2,1e6] so the PSD limiter / breakup / number-adjustment
    regimes are all exercised.  ``frac_empty`` of the points get exact zeros in q_lcl
    and/or q_rai (the eps-gated branches)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    T, p, rho, q_vap, qsl, qsi = atmospheric_profile(n, rng)
    q_lcl = np.maximum(0.0, q_vap - qsl) + rng.random(n) * 1e-4
    q_rai = rng.random(n) * 1e-4
    if number == "const":
        n_lcl = np.full(n, 1e8)
        n_rai = np.full(n, 1e4)
    else:
        n_lcl = 10.0 ** rng.uniform(7, 9, n)
        n_rai = 10.0 ** rng.uniform(2, 6, n)
    if frac_empty > 0:
        u = rng.random(n)
        q_lcl[u < frac_empty] = 0.0
        q_rai[(u > frac_empty / 2) & (u < 1.5 * frac_empty)] = 0.0
        n_rai[(u > 1.5 * frac_empty) & (u < 1.75 * frac_empty)] = 0.0
    q_tot = q_vap + q_lcl + q_rai
    st = dict(rho=rho, T=T, q_tot=q_tot, q_lcl=q_lcl, n_lcl=n_lcl, q_rai=q_rai, n_rai=n_rai)
    return {k: np.ascontiguousarray(v, dtype=dtype) for k, v in st.items()}


def synthetic_states_1m(n, seed=1234, dtype=np.float64, frac_empty=0.05):
    """Inputs of the 1-moment method: rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno
    (test/gpu_performance.jl:80-136)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    T, p, rho, q_vap, qsl, qsi = atmospheric_profile(n, rng)
    q_lcl = np.maximum(0.0, q_vap - qsl) + rng.random(n) * 1e-4
    q_icl = np.maximum(0.0, q_vap - qsi) + rng.random(n) * 1e-4
    q_rai = rng.random(n) * 1e-4
    q_sno = rng.random(n) * 1e-4
    if frac_empty > 0:
        u = rng.random(n)
        q_lcl[u < frac_empty] = 0.0
        q_icl[(u > 0.5 * frac_empty) & (u < 1.5 * frac_empty)] = 0.0
        q_rai[(u > 1.5 * frac_empty) & (u < 2.5 * frac_empty)] = 0.0
        q_sno[(u > 2.5 * frac_empty) & (u < 3.5 * frac_empty)] = 0.0
    q_tot = q_vap + q_lcl + q_icl + q_rai + q_sno
    st = dict(rho=rho, T=T, q_tot=q_tot, q_lcl=q_lcl, q_icl=q_icl, q_rai=q_rai, q_sno=q_sno, p=p)
    return {k: np.ascontiguousarray(v, dtype=dtype) for k, v in st.items()}


def synthetic_states_activation(n, seed=1234, dtype=np.float64, with_hydrometeors=False):
    """Inputs of the ice-nucleation / ARG2000 kernel: T in U[190,300] K,
    p in U[2e4,1.01e5] Pa, w log-uniform in [0.01,10] m/s, q_tot at RH in U[0.5,1.1] over liquid."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d = DEFAULTS
    Rd, Rv = d["gas_constant_dry_air"], d["gas_constant_vapor"]
    T = rng.uniform(190.0, 300.0, n)
    p = rng.uniform(2e4, 1.01e5, n)
    w = 10.0 ** rng.uniform(-2, 1, n)
    RH = rng.uniform(0.5, 1.1, n)
    pv = RH * psat_liq(T)
    eps_ = Rd / Rv
    q_vap = eps_ * pv / (p - (1 - eps_) * pv)
    if with_hydrometeors:
        q_liq, q_ice = rng.random(n) * 1e-4, rng.random(n) * 1e-5
        N_liq, N_ice = np.full(n, 1e3) * 10.0 ** rng.uniform(0, 5, n), np.full(n, 1e3) * 10.0 ** rng.uniform(0, 2, n)
        N_liq[rng.random(n) < 0.1] = 0.0
        N_ice[rng.random(n) < 0.3] = 0.0
    else:
        q_liq = q_ice = N_liq = N_ice = np.zeros(n)
    st = dict(T=T, p=p, w=w, q_tot=q_vap + q_liq + q_ice, q_liq=q_liq, q_ice=q_ice, N_liq=N_liq, N_ice=N_ice)
    return {k: np.ascontiguousarray(v, dtype=dtype) for k, v in st.items()}


def synthetic_states_fused(n, seed=1234, dtype=np.float64):
    """The 11 input columns of the fused 1M+2M+ice-nucleation kernel (config 5): the 1-moment
    atmospheric state plus updraft speed and number concentrations."""
    st = synthetic_states_1m(n, seed=seed, dtype=np.float64)
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    st["w"] = 10.0 ** rng.uniform(-2, 1, n)
    st["n_lcl"] = 10.0 ** rng.uniform(7, 9, n)
    st["n_rai"] = 10.0 ** rng.uniform(2, 6, n)
    keys = ("rho", "T", "p", "w", "q_tot", "q_lcl", "q_icl", "q_rai", "q_sno", "n_lcl", "n_rai")
    return {k: np.ascontiguousarray(st[k], dtype=dtype) for k in keys}


def arg_test_distribution(kind="kappa"):
    """The three modes of SURVEY.md §8d / test/aerosol_activation_tests.jl:41-54: accumulation
    (sea salt), coarse (sea salt), paper mode (sulfate)."""
    from . import AerosolModel as AM
    seasalt = dict(M=0.058443, nu=2.0, rho=2170.0, phi=0.9, kappa=1.12, eps=1.0)
    sulfate = dict(M=0.132, nu=3.0, rho=1770.0, phi=1.0, kappa=0.53, eps=1.0)
    spec = [(0.243e-6, 1.4, 1e8, seasalt), (1.5e-6, 2.1, 1e6, seasalt), (0.05e-6, 2.0, 1e8, sulfate)]
    modes = []
    for r, s, N, a in spec:
        if kind == "kappa":
            modes.append(AM.Mode_κ(r, s, N, (1.0,), (1.0,), (a["M"],), (a["kappa"],)))
        else:
            modes.append(AM.Mode_B(r, s, N, (1.0,), (a["eps"],), (a["phi"],), (a["M"],), (a["nu"],), (a["rho"],)))
    return AM.AerosolDistribution(tuple(modes))


def compare_report(got, ref, bound=None, rtol=F64_RTOL, bound_factor=2.0):
    """Parity metrics of one Float64 output column.

    A point passes when |got - ref| <= rtol * |ref| (the north-star 1e-12 relative
    criterion).  A point that fails it is *excused* only if ``bound`` is given and
    |got - ref| <= bound_factor * bound, where ``bound`` is the first-order rounding-error bound of the
    REFERENCE ALGORITHM ITSELF at that point (oracle_tracked.hpp): the reference subtracts nearly equal
    numbers there and two correct Float64 implementations cannot agree more closely.  bound_factor = 2
    (the reference's own rounding plus ours; measured worst case 0.90).
    Everything else is ``n_bad``.  Exact zeros (gated-off regimes) and non-finite values
    must coincide exactly (``n_zero_mismatch``, ``n_nonfinite_mismatch``)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape
    fin = np.isfinite(ref)
    nonfinite_mismatch = int(np.sum((~fin) & ~((got == ref) | (np.isnan(got) & np.isnan(ref)))))
    nonfinite_mismatch += int(np.sum(fin & ~np.isfinite(got)))
    both = fin & np.isfinite(got)
    diff = np.zeros_like(ref)
    diff[both] = np.abs(got[both] - ref[both])
    denom = np.maximum(np.abs(ref), np.finfo(np.float64).tiny)
    rel = np.where(both, diff / denom, 0.0)
    rel[both & (diff == 0)] = 0.0
    fwd_ok = rel <= rtol
    excused = np.zeros_like(fwd_ok)
    ratio = 0.0
    if bound is not None:
        bnd = np.abs(np.asarray(bound, dtype=np.float64))
        excused |= both & ~fwd_ok & (diff <= bound_factor * bnd)
        need = both & ~fwd_ok & (bnd > 0)
        if need.any():
            ratio = float(np.max(diff[need] / bnd[need]))
    bad = both & ~fwd_ok & ~excused
    zero_mismatch = int(np.sum(both & ((ref == 0) != (got == 0))))
    counted = both & ~excused
    return dict(
        n=int(ref.size),
        max_rel=float(rel[counted].max()) if counted.any() else 0.0,
        max_rel_all=float(rel.max()) if rel.size else 0.0,
        n_excused=int(excused.sum()),
        max_diff_over_bound=ratio,
        n_bad=int(bad.sum()),
        n_nonfinite_mismatch=nonfinite_mismatch,
        n_zero_mismatch=zero_mismatch,
        frac_forward_ok=float(fwd_ok[both].mean()) if both.any() else 1.0,
        worst_index=int(np.argmax(np.where(bad, rel, -1.0))) if bad.any() else (int(np.argmax(np.where(counted, rel, -1.0))) if rel.size else -1),
    )


def assert_parity(name, got, ref, bound=None, **kw):
    """Raise AssertionError with the full report unless the column passes."""
    rep = compare_report(got, ref, bound=bound, **kw)
    ok = rep["n_bad"] == 0 and rep["n_zero_mismatch"] == 0 and rep["n_nonfinite_mismatch"] == 0
    assert ok, (name, rep)
    return rep


def ulp_error_f32(got, truth64):
    """|got - truth| in units of the Float32 ULP of the true value."""
    got = np.asarray(got, dtype=np.float64)
    truth64 = np.asarray(truth64, dtype=np.float64)
    t32 = np.abs(truth64.astype(np.float32))
    ulp = np.spacing(np.maximum(t32, np.finfo(np.float32).tiny)).astype(np.float64)
    return np.abs(got - truth64) / ulp


F32_REPORT = []   # one record per assert_f32_method call with a genuine Float32 reference (tests dump it, tools/parity_report.py prints it)


def assert_f32_method(name, got32, ref32, truth64, bound64=None, max_ulps=F32_ULPS, ref_is_f32_oracle=False):
    """Float32 method criterion: every output within ``max_ulps`` Float32
    ULPs of the true value (the Float64 reference evaluated on the same Float32 inputs and
    parameters), except where the reference algorithm's own Float64 rounding bound already
    exceeds one Float32 ULP; exact zeros / non-finite values coincide with the Float32
    reference's (its regime selection).

    ``ref_is_f32_oracle``: ``ref32`` is the output of the oracle's Float32 instantiation (the reference's own Float32
    arithmetic, ``orc<float>``).  The ULP distance of the GPU result to it — and of it to the true value — is then recorded
    in ``F32_REPORT``: the north star words the Float32 tolerance against the reference's Float32 implementation, whose own
    rounding error (not ours) dominates that distance."""
    got32 = np.asarray(got32)
    assert got32.dtype == np.float32, name
    ref32 = np.asarray(ref32)
    fin = np.isfinite(ref32)
    assert np.array_equal(np.isfinite(got32), fin), (name, "non-finite pattern differs")
    assert np.array_equal((got32 == 0)[fin], (ref32 == 0)[fin]), (name, "zero (regime) pattern differs",
                                                                     int(np.sum((got32 == 0)[fin] != (ref32 == 0)[fin])))
    err = ulp_error_f32(got32[fin], np.asarray(truth64)[fin])
    if bound64 is not None:
        t32 = np.abs(np.asarray(truth64, dtype=np.float64)[fin].astype(np.float32))
        ulp = np.spacing(np.maximum(t32, np.finfo(np.float32).tiny)).astype(np.float64)
        err = np.where(2 * np.abs(np.asarray(bound64)[fin]) > ulp, 0.0, err)
    worst = float(err.max()) if err.size else 0.0
    if ref_is_f32_oracle and err.size:
        t = np.asarray(truth64)[fin]
        d_ref = ulp_error_f32(got32[fin], ref32[fin].astype(np.float64))        # GPU vs the reference's Float32 arithmetic
        e_ref = ulp_error_f32(ref32[fin], t)                                    # the reference's Float32 arithmetic vs the truth
        nz = (ref32[fin] != 0)
        q = lambda a: [float(np.percentile(a[nz], p)) for p in (50, 99)] + [float(a[nz].max())] if nz.any() else [0.0, 0.0, 0.0]
        F32_REPORT.append(dict(name=name, n=int(nz.sum()), gpu_vs_truth_ulp_p50_p99_max=q(err), gpu_vs_ref32_ulp_p50_p99_max=q(d_ref),
                               ref32_vs_truth_ulp_p50_p99_max=q(e_ref),
                               frac_gpu_within_4ulp_of_ref32=float(np.mean(d_ref[nz] <= 4)) if nz.any() else 1.0,
                               frac_gpu_closer_to_truth_than_ref32=float(np.mean(err[nz] <= e_ref[nz])) if nz.any() else 1.0))
    assert worst <= max_ulps, (name, "max Float32 ULP error", worst, "at", int(np.argmax(err)))
    return worst


def synthetic_states_p3(n, seed=1234, dtype=np.float64, frac_ice_free=0.3, frac_unrimed=0.15):
    """Inputs of the 2-moment + P3 method (BMT:898-1083) without logλ: the 2-moment warm-rain state plus
    q_ice, n_ice, q_rim, b_rim.  Follows test/gpu_performance.jl:244-248 (ice content from the profile)
    with a mixed population: ~30 % ice-free points (the BMT:961 branch), unrimed
    (F_rim = 0), partially and heavily rimed ice, rime densities 100-900 kg/m³, mean particle masses
    1e-12..1e-7 kg.  logλ comes from get_distribution_logλ_from_prognostic (host model does the same)."""
    st = synthetic_states_2m(n, seed=seed, dtype=np.float64)
    rng = np.random.Generator(np.random.PCG64(seed + 77))
    T, rho = st["T"], st["rho"]
    qsi = psat_ice(T) / (rho * DEFAULTS["gas_constant_vapor"] * T)
    q_vap = st["q_tot"] - st["q_lcl"] - st["q_rai"]
    q_ice = np.maximum(0.0, q_vap - qsi) * rng.random(n) + rng.random(n) * 1e-4
    x_mean = 10.0 ** rng.uniform(-12, -7, n)
    n_ice = q_ice / x_mean
    F_rim = rng.uniform(0.0, 0.95, n)
    F_rim[rng.random(n) < frac_unrimed] = 0.0
    rho_rim = rng.uniform(100.0, 900.0, n)
    q_rim = F_rim * q_ice
    b_rim = q_rim / rho_rim
    u = rng.random(n)
    q_ice[u < frac_ice_free * 0.6] = 0.0
    n_ice[(u > frac_ice_free * 0.5) & (u < frac_ice_free)] = 0.0
    q_rim = np.minimum(q_rim, q_ice)
    st["q_tot"] = st["q_tot"] + q_ice
    st.update(q_ice=q_ice, n_ice=n_ice, q_rim=q_rim, b_rim=b_rim)
    return {k: np.ascontiguousarray(v, dtype=dtype) for k, v in st.items()}


def train_arg_emulator(orc, n_train=3000, hidden=(32, 16), seed=7, activation="relu", target_transform=True, kind="kappa"):
    """A machine for the emulator methods the way the reference's pipeline makes one (ext/Common.jl): rows of
    [mode features (modes 1 and i swapped), velocity, initial_temperature, initial_pressure], log-preprocessed and standardized,
    fitted to the (target-transformed) ARG2000 activated fraction of the first listed mode — here with scikit-learn's
    MLPRegressor standing in for the MLJ model.  Returns (EmulatorMLP, predict) with ``predict(X)`` scikit-learn's own evaluation
    of the raw feature table, for the tests to compare against.  TEST INFRASTRUCTURE: needs the CPU oracle for the targets."""
    from sklearn.neural_network import MLPRegressor
    from sklearn.preprocessing import StandardScaler
    from oracle import emulator as oe
    from . import AerosolActivation as AA, parameters as CMP
    from .EmulatorModels import EmulatorMLP
    ad = arg_test_distribution(kind)
    ap = CMP.AerosolActivationParameters(np.float64)
    tps = CMP.ThermodynamicsParameters(np.float64)
    hyg = [float(h) for h in AA.mean_hygroscopicity_parameter(ap, ad)]
    st = synthetic_states_activation(n_train, seed=seed)
    blk = CMP.pack_icenuc(tps, ad=ad)
    out = orc.arg_icenuc(blk, *[st[k] for k in ("T", "p", "w", "q_tot", "q_liq", "q_ice", "N_liq", "N_ice")])
    nm = len(ad.modes)
    X, y = [], []
    for i in range(nm):
        sel = slice(i, None, nm)
        X.append(oe.feature_rows(ad.modes, hyg, st["T"][sel], st["p"][sel], st["w"][sel], i))
        y.append(out["N_act"][i][sel] / ad.modes[i].N)
    X, y = np.concatenate(X), np.clip(np.concatenate(y), 0.0, 1.0)
    Xp = oe.preprocess(X, nm)
    scaler = StandardScaler().fit(Xp)
    scaler.scale_ = np.where(scaler.scale_ < 1e-300, 1.0, scaler.scale_)   # constant columns (the spectator modes' kappa ...)
    yt = np.arctanh(2.0 * 0.99 * (y - 0.5)) if target_transform else y      # ext/Common.jl:154-156
    mlp = MLPRegressor(hidden_layer_sizes=hidden, activation=activation, max_iter=400, random_state=seed, tol=1e-6).fit(scaler.transform(Xp), yt)
    machine = EmulatorMLP.from_sklearn(mlp, scaler, log_features=True, target_transform=target_transform)

    def predict(Xraw):
        yy = mlp.predict(scaler.transform(oe.preprocess(Xraw, nm)))
        return oe.inverse_target_transform(yy) if target_transform else yy
    return machine, predict, ad, ap, tps, hyg
