#!/usr/bin/env python
"""Chunk-size sweep of the host-buffer pipeline (cumicro_bmt2m_warm_host_f64): points/s end to end, PCIe inside the timed region."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cumicro  # noqa: E402,F401
from cumicro import BMT, CMP  # noqa: E402
from cumicro.testing import synthetic_states_2m  # noqa: E402

n = 1 << 24
st = synthetic_states_2m(n)
KEYS = ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai")
pinned = [torch.from_numpy(st[k]).pin_memory() for k in KEYS]
out = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(4)]
mp, tps, scheme = CMP.Microphysics2MParams(np.float64), CMP.ThermodynamicsParameters(np.float64), BMT.Microphysics2Moment()
for lg in [int(a) for a in sys.argv[1:]] or [18, 19, 20, 21, 22]:
    f = lambda: BMT.bulk_microphysics_tendencies_host(scheme, mp, tps, *pinned, out=out, chunk=1 << lg)
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        f()
    dt = (time.perf_counter() - t0) / 5
    print(f"chunk 2^{lg}: {dt * 1e3:.2f} ms  {n / dt:.3e} points/s  ({(7 + 4) * 8 * n / dt / 1e9:.1f} GB/s both directions)", flush=True)
