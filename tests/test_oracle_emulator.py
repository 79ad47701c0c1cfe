"""The trained-emulator methods (ext/EmulatorModelsExt.jl) on the CPU: the oracle's restatement of the extension and of the
multilayer-perceptron pipeline against scikit-learn's own evaluation of a fitted model, the feature-row layout, and the C-ABI's
argument checks.  No GPU."""
import ctypes as C

import numpy as np
import pytest

pytest.importorskip("sklearn")


@pytest.fixture(scope="module")
def trained(built, orc):
    from cumicro.testing import train_arg_emulator
    return train_arg_emulator(orc)


def test_feature_rows_swap_the_first_and_the_requested_mode(built):
    """EmulatorModelsExt.jl:47-66: columns mode_1_*, ..., velocity, initial_temperature, initial_pressure; :50-51 swaps 1 <-> i."""
    from oracle import emulator as oe
    from cumicro.testing import arg_test_distribution
    ad = arg_test_distribution("kappa")
    hyg = [0.1, 0.2, 0.3]
    T, p, w = np.array([280.0, 250.0]), np.array([9e4, 5e4]), np.array([0.5, 2.0])
    X = oe.feature_rows(ad.modes, hyg, T, p, w, 2)
    assert X.shape == (2, 15)
    m1, m2, m3 = ad.modes
    assert list(X[0, :4]) == [m3.N, m3.r_dry, m3.stdev, 0.3] and list(X[0, 4:8]) == [m2.N, m2.r_dry, m2.stdev, 0.2]
    assert list(X[0, 8:12]) == [m1.N, m1.r_dry, m1.stdev, 0.1]
    assert list(X[1, 12:]) == [2.0, 250.0, 5e4]
    X0 = oe.feature_rows(ad.modes, hyg, T, p, w, 0)
    assert list(X0[0, :4]) == [m1.N, m1.r_dry, m1.stdev, 0.1]


def test_oracle_pipeline_equals_sklearn_predict(trained):
    """The oracle's plain-Float64 MLP pipeline against scikit-learn's evaluation of the same fitted model."""
    from oracle import emulator as oe
    from cumicro.testing import synthetic_states_activation
    machine, predict, ad, ap, tps, hyg = trained
    st = synthetic_states_activation(500, seed=99)
    for i in range(3):
        X = oe.feature_rows(ad.modes, hyg, st["T"], st["p"], st["w"], i)
        mine = oe.mlp_predict(machine.layers, machine.activation, X, True, machine.feat_mean, machine.feat_scale, machine.target_transform)
        ref = predict(X)
        assert np.max(np.abs(mine - ref)) < 1e-12
    # the emulator does emulate: its activated fraction follows ARG2000's on fresh states (loose: a 400-iteration fit)
    from cumicro import parameters as CMP
    from oracle import oracle as orc
    out = orc.arg_icenuc(CMP.pack_icenuc(tps, ad=ad), *[st[k] for k in ("T", "p", "w", "q_tot", "q_liq", "q_ice", "N_liq", "N_ice")])
    emu = oe.N_activated_per_mode(predict, ad.modes, hyg, st["T"], st["p"], st["w"])
    for i in range(3):
        assert np.mean(np.abs(emu[i] - out["N_act"][i])) / ad.modes[i].N < 0.1
        assert np.all((emu[i] >= 0) & (emu[i] <= ad.modes[i].N))
    tot = oe.total_N_activated(predict, ad.modes, hyg, st["T"], st["p"], st["w"])
    assert np.array_equal(tot, (emu[0] + emu[1]) + emu[2])


def test_weight_count_and_argument_checks_without_a_gpu(built, trained):
    machine, _, ad, ap, _, _ = trained
    abi = built._abi
    lib = abi.load()
    for suf, ft in (("f64", np.float64), ("f32", np.float32)):
        blk = machine.block(ap, ad, suf)
        count = getattr(lib, f"cumicro_emulator_weight_count_{suf}")
        count.restype = C.c_int64
        assert count(C.byref(blk)) == machine.packed(ft).size == 15 * 32 + 32 + 32 * 16 + 16 + 16 + 1
        fn = getattr(lib, f"cumicro_aa_emulated_{suf}")
        null, fake = C.c_void_p(0), C.c_void_p(64)
        assert fn(C.byref(blk), fake, C.c_int64(0), null, null, null, None, null, None) == 0            # empty call
        assert fn(C.byref(blk), fake, C.c_int64(-1), fake, fake, fake, None, fake, None) == -2
        assert fn(C.byref(blk), null, C.c_int64(4), fake, fake, fake, None, fake, None) == -1 and b"weight" in lib.cumicro_last_error()
        assert fn(C.byref(blk), fake, C.c_int64(4), fake, null, fake, None, fake, None) == -1
        assert fn(C.byref(blk), fake, C.c_int64(4), fake, fake, fake, None, null, None) == -1 and b"output" in lib.cumicro_last_error()
        bad = blk.copy()
        bad.width[2] = 2
        assert fn(C.byref(bad), fake, C.c_int64(4), fake, fake, fake, None, fake, None) == -3 and b"width 1" in lib.cumicro_last_error()
        bad = blk.copy()
        bad.n_layers = 5
        assert count(C.byref(bad)) == -1
        bad = blk.copy()
        bad.activation = 9
        assert fn(C.byref(bad), fake, C.c_int64(4), fake, fake, fake, None, fake, None) == -3


def test_host_mirror_validates_the_machine(built):
    from cumicro.EmulatorModels import EmulatorMLP
    rng = np.random.default_rng(0)
    with pytest.raises(ValueError, match="input features"):
        EmulatorMLP([(rng.normal(size=(14, 1)), np.zeros(1))])
    with pytest.raises(ValueError, match="one output"):
        EmulatorMLP([(rng.normal(size=(15, 4)), np.zeros(4))])
    with pytest.raises(ValueError, match="chain"):
        EmulatorMLP([(rng.normal(size=(15, 4)), np.zeros(4)), (rng.normal(size=(5, 1)), np.zeros(1))])
    m = EmulatorMLP([(rng.normal(size=(7, 300 - 44)), np.zeros(256)), (rng.normal(size=(256, 1)), np.zeros(1))], activation="tanh")
    assert m.n_modes == 1 and m.packed(np.float32).dtype == np.float32
