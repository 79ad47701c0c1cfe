#!/usr/bin/env python
"""Launch-shape sweep of the fused config-5 kernel (CUMICRO_FUSED_SHAPE): time per 2^24 points and a checksum of the
outputs (every shape must produce identical tendencies; the diagnostics depend on the block size by rounding only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cumicro  # noqa: E402,F401
from cumicro import CMP, fused  # noqa: E402
from cumicro.testing import arg_test_distribution, synthetic_states_fused  # noqa: E402

dev = torch.device("cuda:0")
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
tps = CMP.ThermodynamicsParameters(np.float64)
st = synthetic_states_fused(n)
c = [torch.from_numpy(st[k]).to(dev) for k in fused.IN_NAMES]
mp1, mp2 = CMP.Microphysics1MParams(np.float64), CMP.Microphysics2MParams(np.float64)
blk3 = CMP.pack_icenuc(tps, ad=arg_test_distribution("kappa"), dust=CMP.DustType("Kaolinite"), hom_linear=True)
o = [torch.empty_like(c[0]) for _ in fused.OUT_NAMES]
ref = None
for shape in sys.argv[2:] or ["128x6", "768x1"]:   # the full sweep (with barriers, 256x3, 384x2) is recorded in kernels_fused.cu
    os.environ["CUMICRO_FUSED_SHAPE"] = shape
    run = lambda: fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *c, out=o)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    chk = [float(torch.nan_to_num(x, nan=0.0, posinf=0.0, neginf=0.0).sum()) for x in o]
    same = ref is None or chk == ref
    ref = ref or chk
    print(f"{shape:8s} {e0.elapsed_time(e1) / 10:.3f} ms  identical={same}", flush=True)
