#!/usr/bin/env python
"""Is the headline bench launch-bound or clock-bound?  Times K back-to-back steps through the public Python API with CUDA events
(GPU ms per step) and with the host clock around the enqueue loop (CPU ms per step), with and without bench.py's NVML sampler
thread.  B200, final round-2 tree: the host enqueues a step in 0.03-0.04 ms against 0.40 ms of kernel; the GPU time per step
grows with the length of the run as the chip reaches its power cap (20 steps 0.399 ms, 200 steps 0.414-0.430, 1000 steps 0.446)."""
import os, sys, time, json
sys.path.insert(0, "/root/repo")
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
def gpu_time(k, sampler):
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx = B.ClockSampler(0) if sampler else None
    if ctx: ctx.__enter__()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(k): step()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    if ctx: ctx.__exit__(None, None, None)
    return e0.elapsed_time(e1) / k, (t1 - t0) / k * 1e3
for k in (20, 200, 200, 1000):
    for s in (False, True):
        g, c = gpu_time(k, s)
        print(f"steps {k:5d} sampler {s!s:5}  gpu ms/step {g:.4f}   cpu enqueue ms/step {c:.4f}", flush=True)
