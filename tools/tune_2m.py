#!/usr/bin/env python
"""Launch-shape sweep of the fused 2M kernel (needs a CUMICRO_TUNING build)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, %r)
import cumicro
from cumicro import BMT, CMP
from cumicro.testing import synthetic_states_2m
n = 1 << 24
st = synthetic_states_2m(n, seed=1234)
K = ("rho","T","q_tot","q_lcl","n_lcl","q_rai","n_rai")
dev = torch.device("cuda:0")
cols = [torch.from_numpy(st[k]).to(dev) for k in K]
outs = [torch.empty_like(cols[0]) for _ in range(4)]
mp = CMP.Microphysics2MParams(np.float64); tps = CMP.ThermodynamicsParameters(np.float64)
f = lambda: BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp, tps, *cols, out=outs)
for _ in range(5): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("variant", os.environ.get("CUMICRO_2M_VARIANT"), "scalar", os.environ.get("CUMICRO_FORCE_SCALAR"), "ms %%.4f  pts/s %%.3e  checksum %%.17g" %% (ms, n / ms * 1e3, float(outs[3].double().sum())))
''' % ROOT
for scalar in ("0", "1"):
    for v in range(9):
        env = dict(os.environ, CUMICRO_2M_VARIANT=str(v), CUMICRO_FORCE_SCALAR=scalar)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-400:], flush=True)
