import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import cumicro
from cumicro import CMP, P3
from cumicro.testing import synthetic_states_p3
n = 1 << int(sys.argv[1])
dev = torch.device("cuda:0")
st = synthetic_states_p3(n)
mp3 = CMP.Microphysics2MParams(np.float64, with_ice=True)
tps = CMP.ThermodynamicsParameters(np.float64)
d = {k: torch.from_numpy(v).to(dev) for k, v in st.items()}
vol = [d[k] * d["rho"] for k in ("q_ice", "n_ice", "q_rim", "b_rim")]
logl = P3.get_distribution_logλ_from_prognostic(mp3, tps, *vol, brent_iters=30)
logl = torch.where(torch.isfinite(logl), logl, torch.zeros_like(logl))
KP = ("rho", "T", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice", "q_rim", "b_rim")
import time
for it in range(3):
    torch.cuda.synchronize(); t = time.time()
    P3.process_rates(mp3, tps, *[d[k] for k in KP], logl)
    torch.cuda.synchronize(); dt = time.time() - t
    print(f"p3_rates n=2^{sys.argv[1]}: {dt*1e3:.2f} ms  {n/dt:.3e} points/s", flush=True)
