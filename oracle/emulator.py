"""TEST INFRASTRUCTURE (CPU oracle) — numpy restatement of the reference's trained-emulator methods, ext/EmulatorModelsExt.jl.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product path never does.

Parity unpinned: the reference's own tests for this extension train their machines at test time from a downloaded data set
(test/aerosol_activation_emulators.jl), so there is no golden vector to pin against.  The restatement is anchored instead on
(1) the extension's source, line by line, with the machine left abstract (`predict`, any callable on the feature table), and
(2) scikit-learn's own `predict` of a fitted MLPRegressor pipeline as an independent implementation of the model class the
device path supports (tests/test_oracle_emulator.py)."""
import numpy as np


def feature_rows(ad_modes, hygro, T, p, w, i):
    """EmulatorModelsExt.jl:47-66: the table handed to the machine for mode i (0-based): columns mode_1_N, mode_1_mean,
    mode_1_stdev, mode_1_kappa, ..., velocity, initial_temperature, initial_pressure, with modes 1 and i swapped (:50-51)."""
    n = len(ad_modes)
    perm = list(range(n))
    perm[0], perm[i] = perm[i], perm[0]
    T = np.asarray(T, dtype=np.float64)
    cols = []
    for j in range(n):
        m = ad_modes[perm[j]]
        cols += [np.full(T.shape, float(m.N)), np.full(T.shape, float(m.r_dry)), np.full(T.shape, float(m.stdev)),
                 np.full(T.shape, float(hygro[perm[j]]))]
    cols += [np.asarray(w, dtype=np.float64), T, np.asarray(p, dtype=np.float64)]
    return np.stack(cols, axis=1)


def N_activated_per_mode(predict, ad_modes, hygro, T, p, w):
    """EmulatorModelsExt.jl:32-69 with `predict(X)` standing for MLJ.predict(machine, X): max(0, min(1, prediction)) * N_i (:67)."""
    out = []
    for i in range(len(ad_modes)):
        frac = np.asarray(predict(feature_rows(ad_modes, hygro, T, p, w, i)), dtype=np.float64)
        clamped = np.where(np.isnan(frac), frac, np.maximum(0.0, np.minimum(1.0, frac)))
        out.append(clamped * float(ad_modes[i].N))
    return out


def total_N_activated(predict, ad_modes, hygro, T, p, w):
    """EmulatorModelsExt.jl:89-103."""
    cols = N_activated_per_mode(predict, ad_modes, hygro, T, p, w)
    tot = cols[0].copy()
    for c in cols[1:]:
        tot = tot + c
    return tot


def preprocess(X, n_modes):
    """ext/Common.jl:57-77 preprocess_aerosol_data: log of every mode_j_N, mode_j_mean and of velocity."""
    X = np.array(X, dtype=np.float64)
    for j in range(n_modes):
        X[:, 4 * j] = np.log(X[:, 4 * j])
        X[:, 4 * j + 1] = np.log(X[:, 4 * j + 1])
    X[:, 4 * n_modes] = np.log(X[:, 4 * n_modes])
    return X


def inverse_target_transform(y):
    """ext/Common.jl:158-160."""
    return (1.0 / (2.0 * 0.99)) * np.tanh(y) + 0.5


def mlp_predict(layers, activation, X, log_features=True, feat_mean=None, feat_scale=None, target_transform=False):
    """The model class of the device path as a plain Float64 pipeline: preprocess -> standardize -> dense layers -> (inverse
    target transform).  layers: [(W[in, out], b[out]), ...]."""
    n_modes = (X.shape[1] - 3) // 4
    h = preprocess(X, n_modes) if log_features else np.array(X, dtype=np.float64)
    if feat_mean is not None:
        h = (h - np.asarray(feat_mean, dtype=np.float64)) * (1.0 / np.asarray(feat_scale, dtype=np.float64))
    act = {"relu": lambda v: np.where(np.isnan(v), v, np.maximum(v, 0.0)), "tanh": np.tanh,
           "logistic": lambda v: 1.0 / (1.0 + np.exp(-v)), "identity": lambda v: v}[activation]
    for k, (W, b) in enumerate(layers):
        acc = np.tile(np.asarray(b, dtype=np.float64), (h.shape[0], 1))
        W = np.asarray(W, dtype=np.float64)
        for j in range(W.shape[0]):          # input order, like a plain dot product
            acc = acc + h[:, j:j + 1] * W[j:j + 1, :]
        h = acc if k == len(layers) - 1 else act(acc)
    y = h[:, 0]
    return inverse_target_transform(y) if target_transform else y
