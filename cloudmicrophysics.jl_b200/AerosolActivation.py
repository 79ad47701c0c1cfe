"""Array-level mirror of ``CloudMicrophysics.AerosolActivation`` (``AA``, ARG2000):
src/AerosolActivation.jl.  The state arguments are device columns; parameters are the host
objects of ``CMP`` / ``AerosolModel``.  One fused kernel evaluates the maximum
supersaturation and every mode's activated number (and mass)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi
from . import AerosolModel as AM
from . import parameters as CMP
from ._columns import check_columns, ptr, ptr_table, stream_handle


def coeff_of_curvature(ap, T):
    """AA.coeff_of_curvature (AA:35-40), host-side scalar or column."""
    return 2 * ap.sigma * ap.M_w / ap.rho_w / ap.R / T


def mean_hygroscopicity_parameter(ap, ad):
    """AA.mean_hygroscopicity_parameter (AA:55-95): tuple over modes; mass-weighted B for
    Mode_B, volume-weighted kappa for Mode_κ.  Parameter-only, evaluated on the host in the
    float type of ``ap``."""
    F = np.float64 if type(ap).__name__.endswith("f64") else np.float32
    out = []
    for m in ad.modes:
        if isinstance(m, AM.Mode_B):
            nom = F(0)
            for j in range(AM.n_components(m)):
                nom += F(m.mass_mix_ratio[j]) * F(m.dissoc[j]) * F(m.osmotic_coeff[j]) * F(m.soluble_mass_frac[j]) / F(m.molar_mass[j])
            den = F(0)
            for j in range(AM.n_components(m)):
                den += F(m.mass_mix_ratio[j]) / F(m.aerosol_density[j])
            out.append(nom / den * F(ap.M_w) / F(ap.rho_w))
        else:
            r = F(0)
            for j in range(AM.n_components(m)):
                r += F(m.vol_mix_ratio[j]) * F(m.kappa[j])
            out.append(r)
    return tuple(out)


def critical_supersaturation(ap, ad, T):
    """AA.critical_supersaturation (AA:107-118), host-side (scalar T or column)."""
    A = coeff_of_curvature(ap, T)
    hyg = mean_hygroscopicity_parameter(ap, ad)
    return tuple(2 / h ** 0.5 * (A / 3 / m.r_dry) ** 1.5 for h, m in zip(hyg, ad.modes))


def _run(ap, ad, aip, tps, T, p, w, q_tot, q_liq, q_ice, N_liq=None, N_ice=None, want=("N",), dust=None, koop=None,
         hom_linear=False):
    cols = [T, p, w, q_tot, q_liq, q_ice]
    if N_liq is None:
        N_liq = torch.zeros_like(T)
    if N_ice is None:
        N_ice = torch.zeros_like(T)
    cols += [N_liq, N_ice]
    suf, n, dev = check_columns(cols, ["T", "p", "w", "q_tot", "q_liq", "q_ice", "N_liq", "N_ice"])
    blk = CMP.pack_icenuc(tps, aps=aip, ap=ap, ad=ad, dust=dust, koop=koop, hom_linear=hom_linear)
    nm = AM.n_modes(ad)
    new = lambda: torch.empty_like(T)
    S_max = new() if "S" in want else None
    N_act = [new() for _ in range(nm)] if "N" in want else None
    M_act = [new() for _ in range(nm)] if "M" in want else None
    J = [new() if "J" in want else None for _ in range(3)]
    da_w = new() if "J" in want else None
    counter = torch.zeros(1, dtype=torch.int64, device=dev) if "J" in want else None
    fn = getattr(_abi.load(), f"cumicro_arg_icenuc_{suf}")
    with torch.cuda.device(dev):
        st = fn(C.byref(blk), C.c_int64(n), *[ptr(c) for c in cols], ptr(S_max), ptr_table(N_act) if N_act else None,
                ptr_table(M_act) if M_act else None, ptr(J[0]), ptr(J[1]), ptr(J[2]), ptr(da_w), ptr(counter), stream_handle(dev))
    _abi.check(st, "cumicro_arg_icenuc")
    return dict(S_max=S_max, N_act=N_act, M_act=M_act, J_dep=J[0], J_ABIFM=J[1], J_hom=J[2], da_w=da_w, n_domain_errors=counter)


def max_supersaturation(ap, ad, aip, tps, T, p, w, q_tot, q_liq, q_ice, N_liq=None, N_ice=None):
    """AA.max_supersaturation (AA:138-214)."""
    return _run(ap, ad, aip, tps, T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice, want=("S",))["S_max"]


def N_activated_per_mode(ap, ad, aip, tps, T, p, w, q_tot, q_liq=None, q_ice=None, N_liq=None, N_ice=None):
    """AA.N_activated_per_mode (AA:235-273): tuple of columns, one per mode.  With a trained emulator as the first argument
    (``EmulatorModels.EmulatorMLP``) this is the method the reference's extension adds: ext/EmulatorModelsExt.jl:32-69."""
    from . import EmulatorModels as EM
    if isinstance(ap, EM.EmulatorMLP):   # (machine, ap, ad, aip, tps, T, p, w, qₜ, qₗ, qᵢ): the signature shifted by one
        return EM.N_activated_per_mode(ap, ad, aip, tps, T, p, w, q_tot, q_liq, q_ice, N_liq)
    if q_liq is None or q_ice is None:
        raise TypeError("N_activated_per_mode(ap, ad, aip, tps, T, p, w, q_tot, q_liq, q_ice): q_liq and q_ice are required")
    return tuple(_run(ap, ad, aip, tps, T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice, want=("N",))["N_act"])


def M_activated_per_mode(ap, ad, aip, tps, T, p, w, q_tot, q_liq, q_ice, N_liq=None, N_ice=None):
    """AA.M_activated_per_mode (AA:294-338)."""
    return tuple(_run(ap, ad, aip, tps, T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice, want=("M",))["M_act"])


def total_N_activated(*args, **kw):
    """AA.total_N_activated (AA:355-384): sum over modes in mode order (ext/EmulatorModelsExt.jl:89-103 for an emulator)."""
    from . import EmulatorModels as EM
    if args and isinstance(args[0], EM.EmulatorMLP):
        return EM.total_N_activated(*args, **kw)
    cols = N_activated_per_mode(*args, **kw)
    tot = cols[0].clone()
    for c in cols[1:]:
        tot += c
    return tot


def total_M_activated(*args, **kw):
    """AA.total_M_activated (AA:403-433)."""
    cols = M_activated_per_mode(*args, **kw)
    tot = cols[0].clone()
    for c in cols[1:]:
        tot += c
    return tot


def activation_and_ice_nucleation(ap, ad, aip, tps, dust, koop, T, p, w, q_tot, q_liq, q_ice, N_liq=None, N_ice=None,
                                  hom_linear=False, with_mass=False):
    """BASELINE config 3 in one kernel: S_max, N_act per mode, J_dep, J_ABIFM, J_hom at
    Δa_w = a_w_eT(p_v, T) - a_w_ice(T) of the same state."""
    want = ("S", "N", "J") + (("M",) if with_mass else ())
    return _run(ap, ad, aip, tps, T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice, want=want, dust=dust, koop=koop,
                hom_linear=hom_linear)
