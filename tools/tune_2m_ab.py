#!/usr/bin/env python
"""Round-robin A/B of headline-kernel variants (CUMICRO_TUNING build, CUMICRO_LIB=...): every pass times each variant once
(20 steps) after a pause, so that all variants see the same thermal / power state; prints min and median per variant.
    python tools/tune_2m_ab.py 22,41,44 [passes]"""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cumicro
from cumicro import BMT, CMP
from cumicro.testing import synthetic_states_2m
n = 1 << 24
st = synthetic_states_2m(n, seed=1234)
K = ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai")
dev = torch.device("cuda:0")
cols = [torch.from_numpy(st[k]).to(dev) for k in K]
outs = [torch.empty_like(cols[0]) for _ in range(4)]
mp = CMP.Microphysics2MParams(np.float64); tps = CMP.ThermodynamicsParameters(np.float64)
f = lambda: BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp, tps, *cols, out=outs)
variants = [int(v) for v in sys.argv[1].split(",")]
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 6
res = {v: [] for v in variants}
chk = {}
for p in range(passes):
    for v in (variants if p % 2 == 0 else variants[::-1]):
        os.environ["CUMICRO_2M_VARIANT"] = str(v)
        time.sleep(0.4)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
        res[v].append(e0.elapsed_time(e1) / 20)
        chk[v] = float(sum(o.double().sum() for o in outs))
for v in variants:
    a = sorted(res[v])
    print(json.dumps(dict(variant=v, min=round(a[0], 4), median=round(a[len(a) // 2], 4), max=round(a[-1], 4), checksum=chk[v])), flush=True)
