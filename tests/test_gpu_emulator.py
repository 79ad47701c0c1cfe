"""GPU parity of the trained-emulator methods (ext/EmulatorModelsExt.jl:32-103; cumicro_aa_emulated_*): the device result
against the oracle's restatement of the extension, with the machine evaluated by (a) the oracle's Float64 pipeline and
(b) scikit-learn's own predict of the fitted model."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
pytest.importorskip("sklearn")


@pytest.fixture(scope="module")
def trained(built, orc):
    from cumicro.testing import train_arg_emulator
    return train_arg_emulator(orc)


def _oracle(machine, ad, hyg, T, p, w, dtype=np.float64):
    from oracle import emulator as oe
    layers = [(W.astype(dtype).astype(np.float64), b.astype(dtype).astype(np.float64)) for W, b in machine.layers]
    pred = lambda X: oe.mlp_predict(layers, machine.activation, X, machine.log_features, machine.feat_mean.astype(dtype).astype(np.float64),
                                    1.0 / (1.0 / machine.feat_scale).astype(dtype).astype(np.float64), machine.target_transform)
    return oe.N_activated_per_mode(pred, ad.modes, hyg, T, p, w)


@pytest.mark.parametrize("n", [1, 10, 4097, 1 << 16])
def test_emulated_activation_f64(built, cuda, trained, n):
    from cumicro import AA
    from cumicro.testing import synthetic_states_activation
    from oracle import emulator as oe
    machine, predict, ad, ap, tps, hyg = trained
    st = synthetic_states_activation(n, seed=31 + n)
    d = {k: torch.from_numpy(st[k]).to(cuda) for k in ("T", "p", "w", "q_tot", "q_liq", "q_ice")}
    got = AA.N_activated_per_mode(machine, ap, ad, None, tps, d["T"], d["p"], d["w"], d["q_tot"], d["q_liq"], d["q_ice"])
    ref = _oracle(machine, ad, hyg, st["T"], st["p"], st["w"])
    skl = oe.N_activated_per_mode(predict, ad.modes, hyg, st["T"], st["p"], st["w"])
    assert len(got) == 3
    for i in range(3):
        g = got[i].cpu().numpy()
        assert np.max(np.abs(g - ref[i])) <= 1e-12 * ad.modes[i].N, i          # the fraction is O(1): absolute on the scale of N_i
        assert np.max(np.abs(g - skl[i])) <= 1e-11 * ad.modes[i].N, i          # scikit-learn's own evaluation (BLAS summation order)
        assert np.all((g >= 0) & (g <= ad.modes[i].N))
    tot = AA.total_N_activated(machine, ap, ad, None, tps, d["T"], d["p"], d["w"], d["q_tot"], d["q_liq"], d["q_ice"]).cpu().numpy()
    assert np.array_equal(tot, ((got[0] + got[1]) + got[2]).cpu().numpy())


def test_emulated_activation_f32(built, cuda, trained):
    """Float32 method: Float32 weights and columns, widened exactly, Float64 sums, one rounding — within 1 Float32 ulp of N_i of
    the exact evaluation of the Float32 model."""
    from cumicro import AA
    from cumicro.testing import synthetic_states_activation
    machine, predict, ad, ap, tps, hyg = trained
    st = synthetic_states_activation(20000, seed=5, dtype=np.float32)
    d = {k: torch.from_numpy(st[k]).to(cuda) for k in ("T", "p", "w")}
    got = AA.N_activated_per_mode(machine, ap, ad, None, tps, d["T"], d["p"], d["w"], d["T"], d["T"], d["T"])
    modes32 = [type("M", (), dict(N=float(np.float32(m.N)), r_dry=float(np.float32(m.r_dry)), stdev=float(np.float32(m.stdev)))) for m in ad.modes]
    hyg32 = [float(np.float32(h)) for h in hyg]
    ref = _oracle(machine, type("AD", (), dict(modes=modes32)), hyg32, st["T"].astype(np.float64), st["p"].astype(np.float64),
                  st["w"].astype(np.float64), dtype=np.float32)
    for i in range(3):
        g = got[i].cpu().numpy()
        assert g.dtype == np.float32
        assert np.max(np.abs(g.astype(np.float64) - ref[i])) <= 1.0 * np.spacing(np.float32(modes32[i].N)), i


@pytest.mark.parametrize("nm,widths,act,tt", [(1, (250, 50, 5, 1), "relu", True), (8, (64, 1), "tanh", False), (2, (1,), "identity", False),
                                              (5, (33, 17, 1), "logistic", True), (3, (256, 256, 256, 1), "relu", False)])
def test_model_shapes_and_activations(built, cuda, nm, widths, act, tt):
    """Random machines over the supported shapes (incl. the reference docs' 250-50-5-1 network) against the oracle pipeline."""
    from cumicro import AA, AerosolModel as AM, parameters as CMP
    from cumicro.EmulatorModels import EmulatorMLP
    from cumicro.testing import synthetic_states_activation
    rng = np.random.default_rng(nm * 100 + len(widths))
    k, layers = 4 * nm + 3, []
    for h in widths:
        layers.append((rng.normal(size=(k, h)) / np.sqrt(k), rng.normal(size=h) * 0.1))
        k = h
    machine = EmulatorMLP(layers, activation=act, log_features=True, feat_mean=rng.normal(size=4 * nm + 3), feat_scale=rng.uniform(0.5, 30, 4 * nm + 3),
                          target_transform=tt)
    ad = AM.AerosolDistribution(tuple(AM.Mode_κ(10 ** rng.uniform(-8, -6), rng.uniform(1.3, 2.5), 10 ** rng.uniform(6, 9), (1.0,), (1.0,), (0.1,),
                                               (rng.uniform(0.1, 1.2),)) for _ in range(nm)))
    ap, tps = CMP.AerosolActivationParameters(np.float64), CMP.ThermodynamicsParameters(np.float64)
    hyg = [float(h) for h in AA.mean_hygroscopicity_parameter(ap, ad)]
    n = 3001
    st = synthetic_states_activation(n, seed=nm)
    st["T"][7] = np.nan                                      # a NaN state gives a NaN fraction (max(0, min(1, NaN)) in the reference)
    d = {k_: torch.from_numpy(st[k_]).to(cuda) for k_ in ("T", "p", "w")}
    got = AA.N_activated_per_mode(machine, ap, ad, None, tps, d["T"], d["p"], d["w"], None, None, None)
    ref = _oracle(machine, ad, hyg, st["T"], st["p"], st["w"])
    for i in range(nm):
        g = got[i].cpu().numpy()
        assert np.isnan(g[7]) and np.isnan(ref[i][7])
        ok = ~np.isnan(ref[i])
        assert np.max(np.abs(g[ok] - ref[i][ok])) <= 1e-12 * ad.modes[i].N * max(1, len(widths)), (i, np.max(np.abs(g[ok] - ref[i][ok])) / ad.modes[i].N)
    tot = AA.total_N_activated(machine, ap, ad, None, tps, d["T"], d["p"], d["w"]).cpu().numpy()
    s = got[0].clone()
    for c in got[1:]:
        s += c
    assert np.array_equal(tot, s.cpu().numpy(), equal_nan=True)


def test_wrong_mode_count_is_refused(built, cuda, trained):
    from cumicro import AA, AerosolModel as AM
    machine, _, ad, ap, tps, _ = trained
    one = AM.AerosolDistribution((ad.modes[0],))
    x = torch.ones(4, dtype=torch.float64, device=cuda)
    with pytest.raises(ValueError, match="modes"):
        AA.N_activated_per_mode(machine, ap, one, None, tps, x, x, x, x, x, x)
