#!/usr/bin/env python
"""What the measurement itself costs in bench.py's timed region: K = 20 back-to-back steps timed with (a) two events only,
(b) an event between every two launches, (c) the in-process NVML sampler thread (2 ms period), (d) both.
B200, final tree: 0.395-0.400 | 0.402-0.403 | 0.397-0.400 | 0.399-0.403 ms per step: an event per launch costs ~7 us per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
import cumicro
from cumicro import BMT, CMP
from cumicro.testing import synthetic_states_2m
n = 1 << 24
dev = torch.device("cuda:0")
st = synthetic_states_2m(n, seed=1234)
cols = [torch.from_numpy(st[k]).to(dev) for k in B.KEYS]
outs = [torch.empty_like(cols[0]) for _ in range(4)]
mp = CMP.Microphysics2MParams(np.float64); tps = CMP.ThermodynamicsParameters(np.float64); scheme = BMT.Microphysics2Moment()
step = lambda: BMT.bulk_microphysics_tendencies(scheme, mp, tps, *cols, out=outs)
def run(k, per_step_events, sampler):
    for _ in range(5): step()
    torch.cuda.synchronize()
    time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
    ctx = B.ClockSampler(0) if sampler else None
    if ctx: ctx.__enter__()
    ev[0].record()
    for i in range(k):
        step()
        if per_step_events or i == k - 1: ev[i + 1].record()
    torch.cuda.synchronize()
    if ctx: ctx.__exit__(None, None, None)
    return ev[0].elapsed_time(ev[k]) / k
for rep in range(3):
    print("two events %.4f | per-step events %.4f | sampler 2 ms %.4f | sampler + per-step events %.4f" %
          (run(20, False, False), run(20, True, False), run(20, False, True), run(20, True, True)), flush=True)
