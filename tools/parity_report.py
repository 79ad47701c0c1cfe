#!/usr/bin/env python
"""Parity report (GPU box): per-field error statistics of the CUDA path vs the CPU
oracle, with the worst points printed in full.  Writes gpurun_out/parity_<tag>.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
import cumicro  # noqa: E402
from cumicro import BMT, CM2, CMP  # noqa: E402
from cumicro.testing import compare_report, synthetic_states_2m  # noqa: E402
from oracle import oracle as orc  # noqa: E402

KEYS = ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai")
OUTS = ("dq_lcl_dt", "dn_lcl_dt", "dq_rai_dt", "dn_rai_dt")
dev = torch.device("cuda:0")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
report = {}


def show(name, got, ref, sens, st):
    rep = compare_report(got, ref, bound=sens)
    report[name] = rep
    flag = "OK " if rep["n_bad"] == 0 and rep["n_zero_mismatch"] == 0 and rep["n_nonfinite_mismatch"] == 0 else "BAD"
    print(f"{flag} {name:34s} max_rel={rep['max_rel']:.2e} all={rep['max_rel_all']:.2e} excused={rep['n_excused']} d/b={rep['max_diff_over_bound']:.2f} "
          f"bad={rep['n_bad']} zero_mm={rep['n_zero_mismatch']} nonfin={rep['n_nonfinite_mismatch']} fwd_ok={rep['frac_forward_ok']:.5f}")
    if flag == "BAD":
        i = rep["worst_index"]
        zm = np.nonzero((ref == 0) != (got == 0))[0]
        for j in ([i] + list(zm[:2])):
            print("     idx", j, "got", repr(float(got[j])), "ref", repr(float(ref[j])), "sens", float(sens[j]) if sens is not None else None,
                  {k: repr(float(st[k][j])) for k in st})


for limited in (True, False):
    for number in ("loguniform", "const"):
        n = 1 << 18
        st = synthetic_states_2m(n, seed=1234, number=number)
        mp = CMP.Microphysics2MParams(np.float64, is_limited=limited)
        tps = CMP.ThermodynamicsParameters(np.float64)
        block = CMP.pack_2m_warm(mp, tps)
        cols = {k: torch.from_numpy(v).to(dev) for k, v in st.items()}
        out = BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp, tps, *[cols[k] for k in KEYS])
        leaves = CM2.sb2006_process_rates(mp, tps, *[cols[k] for k in KEYS])
        ref = orc.bmt2m_warm(block, *[st[k] for k in KEYS], leaves=True)
        bnd = orc.bmt2m_warm_bound(block, *[st[k] for k in KEYS], leaves=True)
        sens = dict(bnd)
        for i_, a_ in enumerate(bnd["leaves"]):
            sens[("leaves", i_)] = a_
        print(f"--- 2M warm f64 limited={limited} number={number}")
        for k in OUTS:
            show(f"{k}[{limited},{number}]", out[k].cpu().numpy(), ref[k], sens[k], st)
        for i, nm in enumerate(cumicro._abi.SB2006_LEAVES):
            show(f"leaf:{nm}[{limited},{number}]", leaves[nm].cpu().numpy(), ref["leaves"][i], sens[("leaves", i)], st)
        # terminal velocities
        sb = mp.warm_rain.seifert_beheng
        N_rai, N_lcl = st["n_rai"] * st["rho"], st["n_lcl"] * st["rho"]
        dNr, dNl = cols["n_rai"] * cols["rho"], cols["n_lcl"] * cols["rho"]
        stv = dict(q=st["q_rai"], rho=st["rho"], N=N_rai)
        for nm, vel, ofn in (("rain_sb", CMP.SB2006VelType(np.float64), orc.termvel_2m_rain_sb),
                             ("rain_chen", CMP.Chen2022VelTypeRain(np.float64), orc.termvel_2m_rain_chen)):
            got = CM2.rain_terminal_velocity(sb, vel, cols["q_rai"], cols["rho"], dNr)
            r = dict(zip(("v0", "v1"), ofn(sb.pdf_r, vel, stv["q"], stv["rho"], stv["N"])))
            sv = dict(zip(("v0", "v1"), orc.termvel_bound("termvel_2m_" + nm, sb.pdf_r, vel, stv["q"], stv["rho"], stv["N"])))
            for j, k in enumerate(("v0", "v1")):
                show(f"vel:{nm}.{k}[{limited},{number}]", got[j].cpu().numpy(), r[k], sv[k], stv)
        stc = dict(q=st["q_lcl"], rho=st["rho"], N=N_lcl)
        vel = CMP.StokesRegimeVelType(np.float64)
        got = CM2.cloud_terminal_velocity(sb.pdf_c, vel, cols["q_lcl"], cols["rho"], dNl)
        r = dict(zip(("v0", "v1"), orc.termvel_2m_cloud(sb.pdf_c, vel, stc["q"], stc["rho"], stc["N"])))
        sv = dict(zip(("v0", "v1"), orc.termvel_bound("termvel_2m_cloud", sb.pdf_c, vel, stc["q"], stc["rho"], stc["N"])))
        for j, k in enumerate(("v0", "v1")):
            show(f"vel:cloud.{k}[{limited},{number}]", got[j].cpu().numpy(), r[k], sv[k], stc)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(report, open(os.path.join(ROOT, "gpurun_out", f"parity_{tag}.json"), "w"), indent=1)
