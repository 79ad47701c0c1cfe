#!/usr/bin/env python
"""Small, ragged-size calls of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitizer_driver.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cumicro  # noqa: E402,F401
from cumicro import AA, BMT, CM2, CMP, IN, P3, fused  # noqa: E402
from cumicro.EmulatorModels import EmulatorMLP  # noqa: E402
from cumicro.testing import (arg_test_distribution, synthetic_states_1m, synthetic_states_2m, synthetic_states_activation,  # noqa: E402
                             synthetic_states_fused, synthetic_states_p3)

dev = torch.device("cuda:0")
tps = CMP.ThermodynamicsParameters(np.float64)
dc = lambda st, keys: [torch.from_numpy(st[k]).to(dev) for k in keys]
K2 = ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai")
done = []

mp2 = CMP.Microphysics2MParams(np.float64)
c = dc(synthetic_states_2m(5001), K2)
BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp2, tps, *c); done.append("2M tile kernel (table)")
BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp2, tps, *[x[1:] for x in c]); done.append("2M tile kernel, misaligned columns")
BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), CMP.Microphysics2MParams(np.float64, overrides={"SB2006_autoconversion_correcting_function_coeff_b": 2.5}), tps, *c)
done.append("2M generic body")
CM2.sb2006_process_rates(mp2, tps, *c); done.append("SB2006 leaves (tiles)")
z = torch.zeros_like(c[0])
CM2.rain_evaporation(mp2, tps, c[2], c[3], z, c[5], z, c[0], c[0] * c[6], c[1]); done.append("rain evaporation leaf (tiles)")

mp1 = CMP.Microphysics1MParams(np.float64)
c1 = dc(synthetic_states_1m(3001), ("rho", "T", "q_tot", "q_lcl", "q_icl", "q_rai", "q_sno"))
m1 = BMT.Microphysics1Moment()
BMT.bulk_microphysics_tendencies(BMT.Instantaneous(), m1, mp1, tps, *c1); done.append("1M Instantaneous (tiles, STD)")
BMT.bulk_microphysics_tendencies(BMT.InstantaneousVerbose(), m1, mp1, tps, *c1); done.append("1M Verbose (tiles)")
BMT.bulk_microphysics_tendencies(BMT.LinearizedAverage(), m1, mp1, tps, *c1, Δt=60.0, nsub=2); done.append("1M LinearizedAverage")

F = np.float32
ca = dc(synthetic_states_activation(4097, dtype=F), ("T", "p", "w", "q_tot", "q_liq", "q_ice", "N_liq", "N_ice"))
AA.activation_and_ice_nucleation(CMP.AerosolActivationParameters(F), arg_test_distribution("kappa"), CMP.AirProperties(F), CMP.ThermodynamicsParameters(F),
                                 CMP.DustType("Kaolinite", F), CMP.Koop2000(F), *ca, hom_linear=False)
done.append("ARG2000 + ice nucleation f32 (tiles 896x1, DomainError counter)")

st = synthetic_states_fused(2003)
cf = dc(st, fused.IN_NAMES)
blk3 = CMP.pack_icenuc(tps, ad=arg_test_distribution("kappa"), dust=CMP.DustType("Kaolinite"), hom_linear=True)
fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *cf); done.append("fused config-5 kernel + diagnostics finish")
from cumicro import collective  # noqa: E402
win = collective.P2PWindow(0, 1)           # single rank: the same exchange code against the rank's own window
buf = torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float64, device=dev)
win.all_reduce(buf); win.all_reduce(buf)
fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *cf, p2p_window=win); done.append("peer-memory exchange window (stand-alone kernel + fused finish kernel, 1 rank)")
torch.cuda.synchronize(); win.destroy()

sp = synthetic_states_p3(700)
d = {k: torch.from_numpy(v).to(dev) for k, v in sp.items()}
mp3 = CMP.Microphysics2MParams(np.float64, with_ice=True)
vol = [d[k] * d["rho"] for k in ("q_ice", "n_ice", "q_rim", "b_rim")]
logl = P3.get_distribution_logλ_from_prognostic(mp3, tps, *vol); done.append("P3 logλ solve")
KP = ("rho", "T", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice", "q_rim", "b_rim")
P3.process_rates(mp3, tps, *[d[k] for k in KP], logl); done.append("P3 process rates")
P3.process_rates(mp3, tps, *[d[k] for k in KP], None); done.append("P3 process rates, logλ solved in the call")
P3.ice_terminal_velocities_from_prognostic(mp3, tps, d["rho"], *vol, None); done.append("P3 velocities")
BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp3, tps, *[d[k] for k in ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice", "q_rim", "b_rim")], logl)
done.append("2M + P3 tendencies")
IN.f23_and_bigg_rates(mp3, tps, *[d[k] for k in ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice")]); done.append("F23 / Bigg rates (tiles)")

ce = dc(synthetic_states_activation(1001), ("T", "p", "w"))
rng = np.random.default_rng(0)
for widths in ((32, 16, 1), (250, 50, 5, 1)):
    k, layers = 15, []
    for h in widths:
        layers.append((rng.normal(size=(k, h)) / np.sqrt(k), rng.normal(size=h) * 0.1))
        k = h
    AA.total_N_activated(EmulatorMLP(layers, activation="relu", target_transform=True), CMP.AerosolActivationParameters(np.float64),
                         arg_test_distribution("kappa"), None, tps, *ce)
done.append("emulator (DMMA dense layers), two networks")
torch.cuda.synchronize()
print("sanitizer driver ran:", "; ".join(done))
