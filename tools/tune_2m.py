#!/usr/bin/env python
"""Launch-shape / variant sweep of the fused 2M kernel (needs a CUMICRO_TUNING build: CUMICRO_TUNING=1 python __graft_entry__.py --force).
One process; the variant is re-read from the environment on every call.  Prints ms per 2^24 points, points/s, and the max relative
difference of each variant's outputs from variant 13 (the round-1 kernel)."""
import os, sys, json
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
variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [13, 14, 0, 20, 21, 22, 23, 24, 30, 31, 32, 33, 34]
waves = sys.argv[2].split(",") if len(sys.argv) > 2 else [""]
ref = None
res = []
for v in variants:
    for wv in waves:
        os.environ["CUMICRO_2M_VARIANT"] = str(v)
        if wv: os.environ["CUMICRO_WAVES"] = wv
        for o in outs: o.zero_()
        for _ in range(5): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for rep in range(3):
            e0.record()
            for _ in range(20): f()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 20)
        o = [x.clone() for x in outs]
        if ref is None: ref = o
        d = max(float(((a - b).abs() / b.abs().clamp_min(1e-300)).max()) for a, b in zip(o, ref))
        nz = sum(int(((a == 0) != (b == 0)).sum()) for a, b in zip(o, ref))
        r = dict(variant=v, waves=wv, ms=round(best, 4), pts_per_s=n / best * 1e3, max_rel_vs_first=d, zero_mismatch=nz)
        res.append(r); print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_2m.json"), "w"), indent=1)
