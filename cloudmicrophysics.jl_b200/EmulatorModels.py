"""Array-level mirror of the reference's package extension ``EmulatorModelsExt`` (ext/EmulatorModelsExt.jl): the methods
``AA.N_activated_per_mode(machine, ap, ad, aip, tps, T, p, w, qₜ, qₗ, qᵢ)`` and ``AA.total_N_activated(machine, ...)`` whose
first argument is a trained emulator of the activated fraction.

The reference dispatches on ``MLJ.Machine`` — any regression model.  The device path takes the model class its own training
pipeline (ext/Common.jl) builds for this job: a multilayer perceptron behind log-preprocessing of N, mean and velocity
(Common.jl:57-77), an optional standardizer, and the optional inverse target transform (Common.jl:154-160).  ``EmulatorMLP``
holds those pieces; ``from_sklearn`` lifts them out of a fitted ``MLPRegressor`` (+ ``StandardScaler``), which is how the tests
obtain a machine trained by an independent implementation."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi
from . import AerosolModel as AM
from ._columns import check_columns, ptr, ptr_table, stream_handle

_ACT = {"relu": 0, "tanh": 1, "logistic": 2, "identity": 3}
MAX_WIDTH, MAX_LAYERS = 256, 4


class EmulatorMLP:
    """layers: [(W, b), ...] with W of shape (inputs, outputs) — input 0 is the feature row
    [mode_1_N, mode_1_mean, mode_1_stdev, mode_1_kappa, ..., velocity, initial_temperature, initial_pressure]
    (EmulatorModelsExt.jl:47-66) after preprocessing; the last layer has one output, the activated fraction of mode 1."""

    def __init__(self, layers, activation="relu", log_features=True, feat_mean=None, feat_scale=None, target_transform=False):
        if not 1 <= len(layers) <= MAX_LAYERS:
            raise ValueError(f"EmulatorMLP: {len(layers)} layers (1..{MAX_LAYERS})")
        if activation not in _ACT:
            raise ValueError(f"EmulatorMLP: activation {activation!r} (one of {sorted(_ACT)})")
        self.layers = [(np.asarray(W, dtype=np.float64), np.asarray(b, dtype=np.float64).reshape(-1)) for W, b in layers]
        k = self.layers[0][0].shape[0]
        if (k - 3) % 4 or not 7 <= k <= 35:
            raise ValueError(f"EmulatorMLP: {k} input features (4 n_modes + 3, n_modes = 1..8)")
        for W, b in self.layers:
            if W.ndim != 2 or W.shape[0] != k or b.shape[0] != W.shape[1] or not 1 <= W.shape[1] <= MAX_WIDTH:
                raise ValueError("EmulatorMLP: layer shapes must chain, (inputs, outputs) each, outputs <= %d" % MAX_WIDTH)
            k = W.shape[1]
        if k != 1:
            raise ValueError("EmulatorMLP: the last layer must have one output")
        self.n_features = self.layers[0][0].shape[0]
        self.n_modes = (self.n_features - 3) // 4
        self.activation = activation
        self.log_features = bool(log_features)
        self.target_transform = bool(target_transform)
        self.feat_mean = np.zeros(self.n_features) if feat_mean is None else np.asarray(feat_mean, dtype=np.float64)
        self.feat_scale = np.ones(self.n_features) if feat_scale is None else np.asarray(feat_scale, dtype=np.float64)
        self._dev = {}

    @classmethod
    def from_sklearn(cls, mlp, scaler=None, log_features=True, target_transform=False):
        """A fitted ``sklearn.neural_network.MLPRegressor`` (and the ``StandardScaler`` in front of it)."""
        return cls(list(zip(mlp.coefs_, mlp.intercepts_)), activation=mlp.activation, log_features=log_features,
                   feat_mean=None if scaler is None else scaler.mean_, feat_scale=None if scaler is None else scaler.scale_,
                   target_transform=target_transform)

    def packed(self, dtype):
        """The weight buffer of include/cumicro.h: per layer W as [inputs][outputs] then b."""
        return np.concatenate([np.concatenate([W.reshape(-1), b]) for W, b in self.layers]).astype(dtype)

    def weights_on(self, dev, dtype):
        key = (str(dev), np.dtype(dtype).str)
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(self.packed(dtype)).to(dev)
        return self._dev[key]

    def block(self, ap, ad, suf):
        from . import AerosolActivation as AA
        nm = AM.n_modes(ad)
        if nm != self.n_modes:
            raise ValueError(f"the emulator was trained for {self.n_modes} modes, the distribution has {nm}")
        hyg = AA.mean_hygroscopicity_parameter(ap, ad)
        pad = lambda v, n: list(v) + [0.0] * (n - len(v))
        blk = _abi.struct("params_emulator", suf)(
            mode_N=pad([m.N for m in ad.modes], 8), mode_mean=pad([m.r_dry for m in ad.modes], 8),
            mode_stdev=pad([m.stdev for m in ad.modes], 8), mode_kappa=pad([float(h) for h in hyg], 8),
            feat_mean=pad(self.feat_mean, 35), feat_inv_scale=pad(1.0 / self.feat_scale, 35),
            n_modes=nm, n_layers=len(self.layers), width=[int(W.shape[1]) for W, _ in self.layers] + [0] * (4 - len(self.layers)),
            activation=_ACT[self.activation], log_features=int(self.log_features), target_transform=int(self.target_transform))
        return blk


def _run(machine, ap, ad, aip, tps, T, p, w, want_total):
    suf, n, dev = check_columns([T, p, w], ["T", "p", "w"])
    blk = machine.block(ap, ad, suf)
    wts = machine.weights_on(dev, np.float64 if suf == "f64" else np.float32)
    lib = _abi.load()
    count = getattr(lib, f"cumicro_emulator_weight_count_{suf}")
    count.restype = C.c_int64
    if count(C.byref(blk)) != wts.numel():
        raise _abi.CuMicroError("emulator weight buffer does not match the layer widths")
    N_act = [torch.empty_like(T) for _ in range(machine.n_modes)]
    N_tot = torch.empty_like(T) if want_total else None
    with torch.cuda.device(dev):
        st = getattr(lib, f"cumicro_aa_emulated_{suf}")(C.byref(blk), ptr(wts), C.c_int64(n), ptr(T), ptr(p), ptr(w), ptr_table(N_act),
                                                         ptr(N_tot), stream_handle(dev))
    _abi.check(st, "cumicro_aa_emulated")
    return N_act, N_tot


def N_activated_per_mode(machine, ap, ad, aip, tps, T, p, w, q_tot=None, q_liq=None, q_ice=None):
    """EmulatorModelsExt.jl:32-69: tuple of columns, clamp(predict(row with modes 1 and i swapped), 0, 1) * N_i.  The water
    contents are accepted and unused, as in the reference's method."""
    return tuple(_run(machine, ap, ad, aip, tps, T, p, w, False)[0])


def total_N_activated(machine, ap, ad, aip, tps, T, p, w, q_tot=None, q_liq=None, q_ice=None):
    """EmulatorModelsExt.jl:89-103: the sum over modes, in mode order, from the same launch."""
    return _run(machine, ap, ad, aip, tps, T, p, w, True)[1]
