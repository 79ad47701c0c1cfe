#!/usr/bin/env python
"""Device-resident times of the secondary kernels (not in bench_families.py): the generic 2-moment body (non-default structure),
the 15-column SB2006 leaves, the stand-alone P3 logλ solve, the rain-evaporation leaf."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cumicro  # noqa: E402
from cumicro import BMT, CM2, CMP, P3  # noqa: E402
from cumicro.testing import synthetic_states_2m, synthetic_states_p3  # noqa: E402

dev = torch.device("cuda:0")


def timeit(f, reps=10, warm=3):
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


n = 1 << 24
tps = CMP.ThermodynamicsParameters(np.float64)
st = synthetic_states_2m(n)
c = [torch.from_numpy(st[k]).to(dev) for k in ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai")]
o = [torch.empty_like(c[0]) for _ in range(4)]
mpg = CMP.Microphysics2MParams(np.float64, overrides={"SB2006_autoconversion_correcting_function_coeff_b": 2.5})
res = {}
try:
    res["2M generic body (acnv.b = 2.5) 2^24"] = timeit(lambda: BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mpg, tps, *c, out=o))
except Exception as e:  # noqa
    res["2M generic body"] = repr(e)[:200]
mp2 = CMP.Microphysics2MParams(np.float64)
res["SB2006 leaves (15 columns) 2^24"] = timeit(lambda: CM2.sb2006_process_rates(mp2, tps, *c), reps=5)
z = torch.zeros_like(c[0])
res["CM2.rain_evaporation leaf 2^24"] = timeit(lambda: CM2.rain_evaporation(mp2, tps, c[2], c[3], z, c[5], z, c[0], c[0] * c[6], c[1]), reps=5)
n4 = 1 << 22
sp = synthetic_states_p3(n4)
d = {k: torch.from_numpy(v).to(dev) for k, v in sp.items()}
mp3 = CMP.Microphysics2MParams(np.float64, with_ice=True)
vol = [d[k] * d["rho"] for k in ("q_ice", "n_ice", "q_rim", "b_rim")]
from cumicro import IN  # noqa: E402
KF = ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice")
res["F23 + Bigg nucleation rates (7 columns) 2^22"] = timeit(lambda: IN.f23_and_bigg_rates(mp3, tps, *[d[k] for k in KF]), reps=5)
res["P3 logλ solve 2^22"] = timeit(lambda: P3.get_distribution_logλ_from_prognostic(mp3, tps, *vol), reps=5, warm=1)
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}))
