#!/usr/bin/env python
"""Runs ONE streaming family a few times so that `ncu -k regex:<kernel> -c 1 -s 2` can capture it:
    python tools/profile_driver.py fused|arg|linavg|inst1m|warm2m|emu [log2n]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cumicro  # noqa: E402,F401
from cumicro import AA, BMT, CMP, fused  # noqa: E402
from cumicro.testing import (arg_test_distribution, synthetic_states_1m, synthetic_states_2m, synthetic_states_activation,  # noqa: E402
                             synthetic_states_fused)

fam = sys.argv[1]
n = 1 << int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
dev = torch.device("cuda:0")
tps = CMP.ThermodynamicsParameters(np.float64)
dc = lambda st, keys: [torch.from_numpy(st[k]).to(dev) for k in keys]
if fam == "fused":
    c = dc(synthetic_states_fused(n), fused.IN_NAMES)
    mp1, mp2 = CMP.Microphysics1MParams(np.float64), CMP.Microphysics2MParams(np.float64)
    blk3 = CMP.pack_icenuc(tps, ad=arg_test_distribution("kappa"), dust=CMP.DustType("Kaolinite"), hom_linear=True)
    o = [torch.empty_like(c[0]) for _ in fused.OUT_NAMES]
    run = lambda: fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *c, out=o)
elif fam == "emu":   # the trained-emulator methods: the reference docs' 15-250-50-5-1 network, random weights
    from cumicro.EmulatorModels import EmulatorMLP
    c = dc(synthetic_states_activation(n), ("T", "p", "w"))
    rng = np.random.default_rng(0)
    k, layers = 15, []
    for h in (250, 50, 5, 1):
        layers.append((rng.normal(size=(k, h)) / np.sqrt(k), rng.normal(size=h) * 0.1))
        k = h
    mach = EmulatorMLP(layers, activation="relu", target_transform=True)
    ap, ad = CMP.AerosolActivationParameters(np.float64), arg_test_distribution("kappa")
    run = lambda: AA.N_activated_per_mode(mach, ap, ad, None, tps, *c, None, None, None)
elif fam == "arg":
    F = np.float32
    c = dc(synthetic_states_activation(n, dtype=F), ("T", "p", "w", "q_tot", "q_liq", "q_ice", "N_liq", "N_ice"))
    args = (CMP.AerosolActivationParameters(F), arg_test_distribution("kappa"), CMP.AirProperties(F), CMP.ThermodynamicsParameters(F),
            CMP.DustType("Kaolinite", F), CMP.Koop2000(F))
    run = lambda: AA.activation_and_ice_nucleation(*args, *c, hom_linear=True)
elif fam in ("linavg", "inst1m"):
    c = dc(synthetic_states_1m(n), ("rho", "T", "q_tot", "q_lcl", "q_icl", "q_rai", "q_sno"))
    mp1 = CMP.Microphysics1MParams(np.float64)
    o = [torch.empty_like(c[0]) for _ in range(4)]
    m1 = BMT.Microphysics1Moment()
    if fam == "linavg":
        run = lambda: BMT.bulk_microphysics_tendencies(BMT.LinearizedAverage(), m1, mp1, tps, *c, Δt=60.0, nsub=1, out=o)
    else:
        run = lambda: BMT.bulk_microphysics_tendencies(BMT.Instantaneous(), m1, mp1, tps, *c, out=o)
else:
    c = dc(synthetic_states_2m(n), ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai"))
    mp2 = CMP.Microphysics2MParams(np.float64)
    o = [torch.empty_like(c[0]) for _ in range(4)]
    run = lambda: BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp2, tps, *c, out=o)
for _ in range(4):
    run()
torch.cuda.synchronize()
print("ok", fam, n)
