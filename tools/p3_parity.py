"""Exploratory P3 parity report (GPU vs oracle), written to gpurun_out/p3_parity.json."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cumicro
from oracle import oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
qorder = int(sys.argv[2]) if len(sys.argv) > 2 else 16
CMP, T_ = cumicro.CMP, cumicro.testing
P3 = cumicro.P3
mp = CMP.Microphysics2MParams(np.float64, with_ice=True, quadrature_order=qorder)
tps = CMP.ThermodynamicsParameters(np.float64)
blk = cumicro.CMP3.pack_p3(mp, tps)
st = T_.synthetic_states_p3(n)
rho = st["rho"]
vol = [st[k] * rho for k in ("q_ice", "n_ice", "q_rim", "b_rim")]
dev = torch.device("cuda:0")
d = {k: torch.from_numpy(v).to(dev) for k, v in st.items()}
dvol = [torch.from_numpy(v).to(dev) for v in vol]
rep = {}
# logλ
t = time.time(); ref_logl = orc.p3_state(blk, *vol, from_prognostic=True, want=("logl",))["logl"]; t_or = time.time() - t
got_logl = P3.get_distribution_logλ_from_prognostic(mp, tps, *dvol).cpu().numpy()
fin = np.isfinite(ref_logl)
rep["logl"] = dict(n=n, oracle_s=t_or, nonfinite_equal=bool(np.array_equal(np.isfinite(got_logl), fin)),
                   max_abs=float(np.max(np.abs(got_logl[fin] - ref_logl[fin]))), n_gt_1e10=int(np.sum(np.abs(got_logl[fin] - ref_logl[fin]) > 1e-10)))
conv = orc.p3_state(blk, *vol, from_prognostic=True, want=("logl",), logl_iters=40)["logl"]
logl = np.where(np.isfinite(conv), conv, 0.0)
dlogl = torch.from_numpy(logl).to(dev)
# stand-alone rates
torch.cuda.synchronize(); t = time.time()
got = P3.process_rates(mp, tps, d["rho"], d["T"], d["q_lcl"], d["n_lcl"], d["q_rai"], d["n_rai"], d["q_ice"], d["n_ice"], d["q_rim"], d["b_rim"], dlogl)
torch.cuda.synchronize(); rep["gpu_rates_s"] = time.time() - t
ice = (st["q_ice"] > 2.3e-16) & (st["n_ice"] > 2.3e-16)
kw = dict(from_prognostic=True, rho_a=rho, T=st["T"], logl=logl, L_c=st["q_lcl"] * rho, N_c=st["n_lcl"] * rho, L_r=st["q_rai"] * rho, N_r=st["n_rai"] * rho)
want = ("v_n", "v_m", "melt", "selfcol", "src7")
t = time.time(); ref = orc.p3_state(blk, *[v[ice] for v in vol], **{k: v[ice] for k, v in kw.items() if k != "from_prognostic"}, from_prognostic=True, want=want); rep["oracle_rates_s"] = time.time() - t
bnd = orc.p3_state(blk, *[v[ice] for v in vol], **{k: v[ice] for k, v in kw.items() if k != "from_prognostic"}, from_prognostic=True, want=want, bound=True)
names = dict(v_n="v_n", v_m="v_m", melt_dNdt="melt_dN", melt_dLdt="melt_dL", self_collection_dNdt="selfcol", dq_c="dq_c", dq_r="dq_r", dN_c="dN_c",
             dN_r="dN_r", dL_rim="dL_rim", dL_ice="dL_ice", dB_rim="dB_rim")
warm = st["T"][ice] > 273.15
for g, r in names.items():
    gg = got[g].cpu().numpy()
    rr, bb = ref[r].copy(), bnd[r].copy()
    if r.startswith("melt"):
        rr = np.where(warm, rr, 0.0); bb = np.where(warm, bb, 0.0)
    rep[g] = T_.compare_report(gg[ice], rr, bound=bb)
    rep[g]["rel_percentiles"] = [float(x) for x in np.percentile(np.abs(gg[ice] - rr) / np.maximum(np.abs(rr), 1e-300), [50, 90, 99, 100])]
    rep[g]["zero_outside_gate"] = bool(np.all(gg[~ice] == 0)) if g not in ("v_n", "v_m") else None
# BMT
t = time.time(); refb = orc.bmt2m_p3(blk, *[st[k] for k in orc.P3_BMT_IN[:-1]], logl); rep["oracle_bmt_s"] = time.time() - t
bndb = orc.bmt2m_p3(blk, *[st[k] for k in orc.P3_BMT_IN[:-1]], logl, bound=True)
torch.cuda.synchronize(); t = time.time()
gotb = cumicro.BMT.bulk_microphysics_tendencies(cumicro.BMT.Microphysics2Moment(), mp, tps, *[d[k] for k in orc.P3_BMT_IN[:-1]], dlogl)
torch.cuda.synchronize(); rep["gpu_bmt_s"] = time.time() - t
for k in orc.P3_BMT_OUT[:-1]:
    rep["bmt_" + k] = T_.compare_report(gotb[k].cpu().numpy(), refb[k], bound=bndb[k])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rep, open("gpurun_out/p3_parity.json", "w"), indent=1)
for k, v in rep.items():
    print(k, v)
