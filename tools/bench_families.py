#!/usr/bin/env python
"""Device-resident throughput of every kernel family (one JSON line per family) — the per-
config numbers quoted in DESIGN.md §5; bench.py remains the contract benchmark (2M, config 1)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cumicro  # noqa: E402
from cumicro import AA, BMT, CMP, fused  # noqa: E402
from cumicro.testing import (arg_test_distribution, synthetic_states_1m, synthetic_states_2m, synthetic_states_activation,  # noqa: E402
                             synthetic_states_fused)

dev = torch.device("cuda:0")
HBM = 6548.5


def timeit(f, reps=20, warm=3):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, n, ms, bytes_per_point):
    print(json.dumps({"family": name, "points": n, "ms": round(ms, 4), "points_per_s": n / ms * 1e3,
                      "algorithmic_GBps": bytes_per_point * n / ms / 1e6, "frac_of_hbm_copy": bytes_per_point * n / ms / 1e6 / HBM}), flush=True)


def dcols(st, keys, dtype=None):
    return [torch.from_numpy(st[k] if dtype is None else st[k].astype(dtype)).to(dev) for k in keys]


tps = CMP.ThermodynamicsParameters(np.float64)
K2 = ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai")
K1 = ("rho", "T", "q_tot", "q_lcl", "q_icl", "q_rai", "q_sno")

n = 1 << 24
c = dcols(synthetic_states_2m(n), K2)
mp2 = CMP.Microphysics2MParams(np.float64)
o = [torch.empty_like(c[0]) for _ in range(4)]
report("2M warm SB2006 f64 (config 2)", n, timeit(lambda: BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp2, tps, *c, out=o)), 88)
c32 = [x.float() for x in c]
mp2f, tpsf = CMP.Microphysics2MParams(np.float32), CMP.ThermodynamicsParameters(np.float32)
o32 = [torch.empty_like(c32[0]) for _ in range(4)]
report("2M warm SB2006 f32", n, timeit(lambda: BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp2f, tpsf, *c32, out=o32)), 44)
del c, c32, o, o32

c = dcols(synthetic_states_1m(n), K1)
mp1 = CMP.Microphysics1MParams(np.float64)
o = [torch.empty_like(c[0]) for _ in range(4)]
m1 = BMT.Microphysics1Moment()
report("1M Instantaneous f64 2^24", n, timeit(lambda: BMT.bulk_microphysics_tendencies(BMT.Instantaneous(), m1, mp1, tps, *c, out=o)), 88)
report("1M InstantaneousVerbose (4 + 18 columns) f64", n, timeit(lambda: BMT.bulk_microphysics_tendencies(BMT.InstantaneousVerbose(), m1, mp1, tps, *c), reps=10), 56 + 8 * 22)
report("1M LinearizedAverage nsub=1 f64", n, timeit(lambda: BMT.bulk_microphysics_tendencies(BMT.LinearizedAverage(), m1, mp1, tps, *c, Δt=60.0, nsub=1, out=o)), 88)
report("1M LinearizedAverage nsub=3 f64", n, timeit(lambda: BMT.bulk_microphysics_tendencies(BMT.LinearizedAverage(), m1, mp1, tps, *c, Δt=60.0, nsub=3, out=o), reps=10), 88)
n1 = 64 ** 3
c1 = [x[:n1].contiguous() for x in c]
o1 = [torch.empty_like(c1[0]) for _ in range(4)]
report("1M Instantaneous f64 64^3 (config 1)", n1, timeit(lambda: BMT.bulk_microphysics_tendencies(BMT.Instantaneous(), m1, mp1, tps, *c1, out=o1), reps=50), 88)
# the same 50 launches captured in a CUDA graph: at 64^3 points the eager figure is the host's call rate (ctypes + launch), not the kernel
try:
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        BMT.bulk_microphysics_tendencies(BMT.Instantaneous(), m1, mp1, tps, *c1, out=o1)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(50):
                BMT.bulk_microphysics_tendencies(BMT.Instantaneous(), m1, mp1, tps, *c1, out=o1)
    torch.cuda.current_stream().wait_stream(side)
    report("1M Instantaneous f64 64^3 (config 1), 50 launches replayed from a CUDA graph", n1, timeit(g.replay, reps=20) / 50, 88)
except Exception as e:  # noqa
    print(json.dumps({"family": "1M Instantaneous 64^3 CUDA graph", "error": repr(e)[:200]}))
del c, o

n3 = 1 << 25
KA = ("T", "p", "w", "q_tot", "q_liq", "q_ice", "N_liq", "N_ice")
F = np.float32
c = dcols(synthetic_states_activation(n3, dtype=F), KA)
tf = CMP.ThermodynamicsParameters(F)
args = (CMP.AerosolActivationParameters(F), arg_test_distribution("kappa"), CMP.AirProperties(F), tf, CMP.DustType("Kaolinite", F), CMP.Koop2000(F))
report("ice nucleation + ARG2000 3 modes f32 2^25 (config 3)", n3, timeit(lambda: AA.activation_and_ice_nucleation(*args, *c, hom_linear=True), reps=10), 4 * (8 + 1 + 3 + 4))
del c

# ---- the trained-emulator variant of N_activated_per_mode (ext/EmulatorModelsExt.jl): random machines of the reference docs' shape
from cumicro.EmulatorModels import EmulatorMLP  # noqa: E402
ne = 1 << 20
ce = dcols(synthetic_states_activation(ne), ("T", "p", "w"))
ad3 = arg_test_distribution("kappa")
ap64 = CMP.AerosolActivationParameters(np.float64)
rng = np.random.default_rng(0)
for widths in ((32, 16, 1), (250, 50, 5, 1)):
    k, layers = 15, []
    for h in widths:
        layers.append((rng.normal(size=(k, h)) / np.sqrt(k), rng.normal(size=h) * 0.1))
        k = h
    mach = EmulatorMLP(layers, activation="relu", target_transform=True)
    ms = timeit(lambda: AA.N_activated_per_mode(mach, ap64, ad3, None, tps, *ce, None, None, None), reps=5, warm=2)
    flops = 2.0 * sum(W.size for W, _ in layers) * 3 * ne
    report(f"emulated N_activated_per_mode, MLP 15-{'-'.join(map(str, widths))}, 3 modes f64 2^20 ({flops / ms / 1e9:.2f} FP64 TFLOP/s)", ne, ms, 8 * 6)
del ce

n5 = 1 << 24
st = synthetic_states_fused(n5)
c = dcols(st, fused.IN_NAMES)
blk3 = CMP.pack_icenuc(tps, ad=arg_test_distribution("kappa"), dust=CMP.DustType("Kaolinite"), hom_linear=True)
o = [torch.empty_like(c[0]) for _ in fused.OUT_NAMES]
report("fused 1M+2M+ice nucleation(+ARG) f64 2^24 per GPU (config 5)", n5,
       timeit(lambda: fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *c, out=o), reps=10), 176)
del c, o

if os.environ.get("CUMICRO_FAMILIES_SKIP_P3"):   # quick iterations on the streaming families
    sys.exit(0)

# ---- P3 (config 4): stand-alone process rates + the fused 2M+P3 tendencies, Float64, 2^22 points
from cumicro import P3  # noqa: E402
from cumicro.testing import synthetic_states_p3  # noqa: E402

n4 = 1 << (int(os.environ.get("CUMICRO_P3_LOG2N", "22")))
st = synthetic_states_p3(n4)
mp3 = CMP.Microphysics2MParams(np.float64, with_ice=True)
KP = ("rho", "T", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice", "q_rim", "b_rim")
d = {k: torch.from_numpy(v).to(dev) for k, v in st.items()}
vol = [d[k] * d["rho"] for k in ("q_ice", "n_ice", "q_rim", "b_rim")]
ms = timeit(lambda: P3.get_distribution_logλ_from_prognostic(mp3, tps, *vol), reps=5, warm=1)
report("P3 get_distribution_logλ_from_prognostic f64", n4, ms, 40)
logl = P3.get_distribution_logλ_from_prognostic(mp3, tps, *vol, brent_iters=30)
logl = torch.where(torch.isfinite(logl), logl, torch.zeros_like(logl))
ice_frac = float(((d["q_ice"] > 2.3e-16) & (d["n_ice"] > 2.3e-16)).double().mean())
ms = timeit(lambda: P3.ice_terminal_velocities_from_prognostic(mp3, tps, d["rho"], *vol, logl), reps=5, warm=1)
report("P3 ice terminal velocities (number + mass weighted) f64", n4, ms, 64)
ms = timeit(lambda: P3.process_rates(mp3, tps, *[d[k] for k in KP], logl), reps=2, warm=1)
report(f"P3 process rates GL(16) f64 2^{int(np.log2(n4))} (config 4; {ice_frac:.2f} of the points ice-bearing)", n4, ms, 192)
ms = timeit(lambda: P3.process_rates(mp3, tps, *[d[k] for k in KP], None), reps=2, warm=1)
report("P3 process rates with the logλ solve inside the call (logλ column NULL)", n4, ms, 184)
# CPU port of the same integrals on a bounded sample (all host cores), for the GPU/CPU ratio quoted in DESIGN.md
from oracle import oracle as orc  # noqa: E402
orc.set_num_threads(os.cpu_count() or 1)
m = 1 << 12
blk = cumicro.CMP3.pack_p3(mp3, tps)
hs = {k: v[:m] for k, v in st.items()}
hl = logl[:m].cpu().numpy()
import time  # noqa: E402
t0 = time.perf_counter()
orc.p3_state(blk, *[hs[k] * hs["rho"] for k in ("q_ice", "n_ice", "q_rim", "b_rim")], from_prognostic=True, rho_a=hs["rho"], T=hs["T"], logl=hl,
             L_c=hs["q_lcl"] * hs["rho"], N_c=hs["n_lcl"] * hs["rho"], L_r=hs["q_rai"] * hs["rho"], N_r=hs["n_rai"] * hs["rho"],
             want=("v_n", "v_m", "melt", "selfcol", "src7"))
dt = time.perf_counter() - t0
print(json.dumps({"family": "P3 process rates, CPU port (oracle, OpenMP)", "points": m, "ms": dt * 1e3, "points_per_s": m / dt,
                  "threads": orc.num_threads()}), flush=True)
cols = [d[k] for k in ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice", "q_rim", "b_rim")] + [logl]
ms = timeit(lambda: BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp3, tps, *cols), reps=2, warm=1)
report("2M + P3 fused tendencies (BMT:898-1083) f64", n4, ms, 160)
