import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import cumicro
from cumicro import BMT, CMP
from cumicro.testing import synthetic_states_1m
dev = torch.device("cuda:0"); n = 1 << 24
tps = CMP.ThermodynamicsParameters(np.float64)
st = synthetic_states_1m(n)
c = [torch.from_numpy(st[k]).to(dev) for k in ("rho", "T", "q_tot", "q_lcl", "q_icl", "q_rai", "q_sno")]
mp1 = CMP.Microphysics1MParams(np.float64); o = [torch.empty_like(c[0]) for _ in range(4)]; m1 = BMT.Microphysics1Moment()
def t(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
print(os.environ.get("CUMICRO_LIB", "default")[-22:], "inst %.3f" % t(lambda: BMT.bulk_microphysics_tendencies(BMT.Instantaneous(), m1, mp1, tps, *c, out=o)),
      "verbose %.3f" % t(lambda: BMT.bulk_microphysics_tendencies(BMT.InstantaneousVerbose(), m1, mp1, tps, *c)),
      "linavg1 %.3f" % t(lambda: BMT.bulk_microphysics_tendencies(BMT.LinearizedAverage(), m1, mp1, tps, *c, Δt=60.0, nsub=1, out=o)),
      "chk", float(sum(x.sum() for x in o)))
