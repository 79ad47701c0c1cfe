"""ctypes binding of the CPU parity oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False):
    """Compile the oracle with the committed Makefile (g++, -ffp-contract=off, OpenMP)."""
    # always through make (incremental: a no-op when the library is newer than every source / header)
    if force or not os.path.exists(LIB_PATH) or os.path.exists(os.path.join(_HERE, "Makefile")):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int):
    lib().oracle_set_num_threads(int(n))


class f32_thresholds:
    """Context manager: Float64 evaluations use the Float32 method's thresholds, i.e. they
    compute the true value of the Float32 method (regimes of Float32, exact arithmetic)."""

    def __enter__(self):
        lib().oracle_set_f32_thresholds(1)

    def __exit__(self, *a):
        lib().oracle_set_f32_thresholds(0)


def _suf(dtype):
    return {"float64": "f64", "float32": "f32"}[np.dtype(dtype).name]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _cols(arrays, dtype):
    out = [np.ascontiguousarray(a, dtype=dtype) for a in arrays]
    n = out[0].shape[0]
    assert all(a.shape == (n,) for a in out)
    return out, n


def bmt2m_warm(params, rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, leaves=False):
    """BMT:820-854 over arrays. Returns dict of 4 tendencies (+ the SB2006 leaf columns)."""
    from importlib import import_module
    dtype = np.float64 if type(params).__name__.endswith("f64") else np.float32
    (rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai), n = _cols((rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai), dtype)
    names = ("dq_lcl_dt", "dn_lcl_dt", "dq_rai_dt", "dn_rai_dt")
    out = {k: np.empty(n, dtype) for k in names}
    leaf_ptrs = None
    leaf_arrays = None
    if leaves:
        nleaf = 15
        leaf_arrays = [np.empty(n, dtype) for _ in range(nleaf)]
        leaf_ptrs = (C.c_void_p * nleaf)(*[_ptr(a) for a in leaf_arrays])
    fn = getattr(lib(), f"oracle_bmt2m_warm_{_suf(dtype)}")
    st = fn(C.byref(params), C.c_int64(n), *[_ptr(a) for a in (rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai)],
            *[_ptr(out[k]) for k in names], leaf_ptrs)
    assert st == 0
    if leaves:
        out["leaves"] = leaf_arrays
    return out


def rain_evaporation_2m(params, q_tot, q_lcl, q_icl, q_rai, q_sno, rho, N_rai, T):
    """CM2.rain_evaporation + CM2.∂rain_evaporation_∂N_rai_∂q_rai over arrays (CM2:780-853): 4 columns."""
    dtype = np.float64 if type(params).__name__.endswith("f64") else np.float32
    cols, n = _cols((q_tot, q_lcl, q_icl, q_rai, q_sno, rho, N_rai, T), dtype)
    names = ("dNrho_dt", "dq_dt", "dN_rai", "dq_rai")
    out = {k: np.empty(n, dtype) for k in names}
    ip = (C.c_void_p * 8)(*[_ptr(a) for a in cols])
    op = (C.c_void_p * 4)(*[_ptr(out[k]) for k in names])
    assert getattr(lib(), f"oracle_rain_evaporation_2m_{_suf(dtype)}")(C.byref(params), C.c_int64(n), ip, op) == 0
    return out


def bmt2m_warm_bound(params, rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, leaves=False):
    """First-order rounding-error bounds of the reference algorithm (oracle_tracked.hpp)
    for the outputs of ``bmt2m_warm`` on the same Float64 inputs."""
    assert type(params).__name__.endswith("f64")
    (rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai), n = _cols((rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai), np.float64)
    names = ("dq_lcl_dt", "dn_lcl_dt", "dq_rai_dt", "dn_rai_dt")
    out = {k: np.empty(n, np.float64) for k in names}
    b4 = (C.c_void_p * 4)(*[_ptr(out[k]) for k in names])
    leaf_ptrs, leaf_arrays = None, None
    if leaves:
        leaf_arrays = [np.empty(n, np.float64) for _ in range(15)]
        leaf_ptrs = (C.c_void_p * 15)(*[_ptr(a) for a in leaf_arrays])
    st = lib().oracle_bmt2m_warm_bound_f64(C.byref(params), C.c_int64(n),
                                           *[_ptr(a) for a in (rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai)], b4, leaf_ptrs)
    assert st == 0
    if leaves:
        out["leaves"] = leaf_arrays
    return out


def termvel_bound(fname, pdf, vel, q, rho, N):
    """Error bounds for ``termvel_2m_{rain_sb,rain_chen,cloud}`` (Float64)."""
    (q, rho, N), n = _cols((q, rho, N), np.float64)
    b0, b1 = np.empty(n), np.empty(n)
    st = getattr(lib(), f"oracle_{fname}_bound_f64")(C.byref(pdf), C.byref(vel), C.c_int64(n), _ptr(q), _ptr(rho),
                                                     _ptr(N), _ptr(b0), _ptr(b1))
    assert st == 0
    return b0, b1


def _termvel(fname, pdf, vel, q, rho, N):
    dtype = np.float64 if type(pdf).__name__.endswith("f64") else np.float32
    (q, rho, N), n = _cols((q, rho, N), dtype)
    vt0 = np.empty(n, dtype)
    vt1 = np.empty(n, dtype)
    fn = getattr(lib(), f"oracle_{fname}_{_suf(dtype)}")
    st = fn(C.byref(pdf), C.byref(vel), C.c_int64(n), _ptr(q), _ptr(rho), _ptr(N), _ptr(vt0), _ptr(vt1))
    assert st == 0
    return vt0, vt1


def termvel_2m_rain_sb(pdf_r, vel, q, rho, N):
    return _termvel("termvel_2m_rain_sb", pdf_r, vel, q, rho, N)


def termvel_2m_rain_chen(pdf_r, vel, q, rho, N):
    return _termvel("termvel_2m_rain_chen", pdf_r, vel, q, rho, N)


def termvel_2m_cloud(pdf_c, vel, q, rho, N):
    return _termvel("termvel_2m_cloud", pdf_c, vel, q, rho, N)


# ---- 1-moment scheme ---------------------------------------------------------------------
SRC_1M = ("S_phase_change_vap_lcl", "S_phase_change_vap_icl", "S_acnv_lcl_rai", "S_acnv_icl_sno", "S_accr_lcl_rai",
          "S_accr_lcl_sno_cold", "S_accr_lcl_sno_warm", "S_accr_melt_lcl_sno", "S_accr_icl_rai", "S_accr_freeze_icl_rai",
          "S_accr_icl_sno", "S_accr_rai_sno_cold", "S_accr_rai_sno_warm", "S_accr_melt_rai_sno", "S_phase_change_vap_rai",
          "S_phase_change_vap_sno", "S_melt_icl_lcl", "S_melt_sno_rai")
OUT_1M = ("dq_lcl_dt", "dq_icl_dt", "dq_rai_dt", "dq_sno_dt")
MODES_1M = {"instantaneous": 0, "verbose": 1, "linearized_average": 2}


def bmt1m(params, rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno, mode="instantaneous", dt=0.0, nsub=1, bound=False):
    """BMT:505-632 over arrays.  mode in MODES_1M; ``bound=True`` returns the rounding-error
    bounds of the reference algorithm (Float64 only) instead of the values."""
    dtype = np.float64 if type(params).__name__.endswith("f64") else np.float32
    cols, n = _cols((rho, T, q_tot, q_lcl, q_icl, q_rai, q_sno), dtype)
    out = {k: np.empty(n, dtype) for k in OUT_1M}
    o4 = (C.c_void_p * 4)(*[_ptr(out[k]) for k in OUT_1M])
    m = MODES_1M[mode]
    s18 = None
    if m == 1:
        for k in SRC_1M:
            out[k] = np.empty(n, dtype)
        s18 = (C.c_void_p * 18)(*[_ptr(out[k]) for k in SRC_1M])
    if bound:
        assert dtype == np.float64
        fn = lib().oracle_bmt1m_bound_f64
    else:
        fn = getattr(lib(), f"oracle_bmt1m_{_suf(dtype)}")
    cdt = C.c_double(dt) if dtype == np.float64 else C.c_float(dt)
    st = fn(C.byref(params), C.c_int(m), C.c_int64(n), *[_ptr(a) for a in cols], cdt, C.c_int(nsub), o4, s18)
    assert st == 0
    return out


TERMVEL_1M = {"rain_blk1m": 0, "snow_blk1m": 1, "rain_chen": 2, "snow_chen": 3, "cloud_liquid_stokes": 4, "cloud_ice_chen": 5}


def termvel_1m(params, kind, rho, q, vel=None, bound=False):
    dtype = np.float64 if type(params).__name__.endswith("f64") else np.float32
    (rho, q), n = _cols((rho, q), dtype)
    out = np.empty(n, dtype)
    fn = lib().oracle_termvel_1m_bound_f64 if bound else getattr(lib(), f"oracle_termvel_1m_{_suf(dtype)}")
    st = fn(C.byref(params), C.byref(vel) if vel is not None else None, C.c_int(TERMVEL_1M[kind]), C.c_int64(n), _ptr(rho), _ptr(q), _ptr(out))
    assert st == 0
    return out


# ---- ice nucleation / water activity / ARG2000 ----------------------------------------------
ICENUC_WHAT = {"deposition_J": 0, "ABIFM_J": 1, "homogeneous_J_cubic": 2, "homogeneous_J_linear": 3, "a_w_ice": 4, "a_w_eT": 5,
               "a_w_xT": 6, "H2SO4_soln_saturation_vapor_pressure": 7, "P3_deposition_N_i": 8, "INP_concentration_mean": 9,
               "dust_activated_number_fraction": 10}


def icenuc(params, what, x, y=None):
    """Pointwise leaves (same numbering as cumicro_icenuc_*).  Returns (out, n_domain_errors)."""
    dtype = np.float64 if type(params).__name__.endswith("f64") else np.float32
    x = np.ascontiguousarray(x, dtype=dtype)
    y = np.ascontiguousarray(y if y is not None else x, dtype=dtype)
    out = np.empty_like(x)
    fn = getattr(lib(), f"oracle_icenuc_{_suf(dtype)}")
    fn.restype = C.c_int64
    nerr = fn(C.byref(params), C.c_int(ICENUC_WHAT[what]), C.c_int64(x.size), _ptr(x), _ptr(y), _ptr(out))
    return out, int(nerr)


def arg_icenuc(params, T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice, bound=False):
    """ARG2000 (AA:138-324) + nucleation rates at Δa_w = a_w_eT(p_v,T) - a_w_ice(T)."""
    dtype = np.float64 if type(params).__name__.endswith("f64") else np.float32
    cols, n = _cols((T, p, w, q_tot, q_liq, q_ice, N_liq, N_ice), dtype)
    nm = params.n_modes
    new = lambda: np.empty(n, dtype)
    out = dict(S_max=new(), N_act=[new() for _ in range(nm)], M_act=[new() for _ in range(nm)], J_dep=new(), J_ABIFM=new(),
               J_hom=new(), da_w=new())
    na = (C.c_void_p * nm)(*[_ptr(a) for a in out["N_act"]])
    ma = (C.c_void_p * nm)(*[_ptr(a) for a in out["M_act"]])
    fn = lib().oracle_arg_icenuc_bound_f64 if bound else getattr(lib(), f"oracle_arg_icenuc_{_suf(dtype)}")
    fn.restype = C.c_int64
    nerr = fn(C.byref(params), C.c_int64(n), *[_ptr(a) for a in cols], _ptr(out["S_max"]), na, ma, _ptr(out["J_dep"]),
              _ptr(out["J_ABIFM"]), _ptr(out["J_hom"]), _ptr(out["da_w"]))
    out["n_domain_errors"] = int(nerr)
    return out


# ---- P3 ice scheme (oracle_p3.hpp; Float64 parameter block) -----------------------------------
P3_LEAF = {"gamma_inc_P": 0, "gamma_inc_inv": 1, "rime_mass_fraction": 2, "rime_density": 3}
P3_COLL10 = ("QCFRZ", "QCSHD", "NCCOL", "QRFRZ", "QRSHD", "NRCOL", "M_col", "BCCOL", "BRCOL", "wet_M_col")
P3_SRC7 = ("dq_c", "dq_r", "dN_c", "dN_r", "dL_rim", "dL_ice", "dB_rim")
P3_THR = ("rho_g", "D_th", "D_gr", "D_cr", "F_rim")
P3_BMT_IN = ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai", "q_ice", "n_ice", "q_rim", "b_rim", "logl")
P3_BMT_OUT = ("dq_lcl_dt", "dn_lcl_dt", "dq_rai_dt", "dn_rai_dt", "dq_ice_dt", "dn_ice_dt", "dq_rim_dt", "db_rim_dt",
              "dn_lcl_activation_dt")


def p3_leaf(what, x, y, bound=False):
    """UT.gamma_inc / gamma_inc_inv / rime_mass_fraction / rime_density over arrays."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(x)
    st = lib().oracle_p3_leaf_f64(C.c_int(P3_LEAF[what]), C.c_int64(x.size), _ptr(x), _ptr(y), _ptr(out), C.c_int(int(bound)))
    assert st == 0
    return out


def p3_state(params, L_ice, N_ice, c3, c4, *, from_prognostic=False, rho_a=None, T=None, logl=None, L_c=None, N_c=None, L_r=None,
             N_r=None, logl_iters=-1, want=("logl",), bound=False):
    """State-level P3 functions over columns.  (c3, c4) = (F_rim, rho_rim), or (L_rim, B_rim) with
    ``from_prognostic``.  ``logl=None`` solves get_distribution_logλ (``logl_iters`` Brent iterations,
    -1 = the reference's 10 / 8).  ``want`` selects outputs among: thresholds, logl, D_m, v_n, v_m, melt,
    selfcol, coll10, src7, max_freeze, rime_local."""
    assert type(params).__name__.endswith("p3_f64")
    req = [L_ice, N_ice, c3, c4]
    (L_ice, N_ice, c3, c4), n = _cols(req, np.float64)
    opt = {}
    for k, v in dict(rho_a=rho_a, T=T, logl=logl, L_c=L_c, N_c=N_c, L_r=L_r, N_r=N_r).items():
        opt[k] = None if v is None else np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (n,)))
    new = lambda: np.empty(n, np.float64)
    out = {}
    thr = None
    if "thresholds" in want:
        for k in P3_THR:
            out[k] = new()
        thr = (C.c_void_p * 5)(*[_ptr(out[k]) for k in P3_THR])
    for k in ("logl", "D_m", "v_n", "v_m", "selfcol", "max_freeze", "rime_local"):
        if k in want:
            out[k] = new()
    if "melt" in want:
        out["melt_dN"], out["melt_dL"] = new(), new()
    c10 = s7 = None
    if "coll10" in want:
        for k in P3_COLL10:
            out[k] = new()
        c10 = (C.c_void_p * 10)(*[_ptr(out[k]) for k in P3_COLL10])
    if "src7" in want:
        for k in P3_SRC7:
            out[k] = new()
        s7 = (C.c_void_p * 7)(*[_ptr(out[k]) for k in P3_SRC7])
    if ("coll10" in want or "src7" in want) and any(opt[k] is None for k in ("L_c", "N_c", "L_r", "N_r")):
        raise ValueError("collisions need L_c, N_c, L_r, N_r")
    g = lambda k: _ptr(out[k]) if k in out else None
    st = lib().oracle_p3_state_f64(
        C.byref(params), C.c_int64(n), C.c_int(int(from_prognostic)), _ptr(L_ice), _ptr(N_ice), _ptr(c3), _ptr(c4),
        _ptr(opt["rho_a"]), _ptr(opt["T"]), _ptr(opt["logl"]), _ptr(opt["L_c"]), _ptr(opt["N_c"]), _ptr(opt["L_r"]), _ptr(opt["N_r"]),
        C.c_int(logl_iters), thr, g("logl"), g("D_m"), g("v_n"), g("v_m"), g("melt_dN"), g("melt_dL"), g("selfcol"), c10, s7,
        g("max_freeze"), g("rime_local"), C.c_int(int(bound)))
    assert st == 0
    return out


def bmt2m_p3(params, rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, q_rim, b_rim, logl, inpc_log_shift=None, bound=False):
    """BMT:898-1083 over arrays -> dict of the 9 tendencies."""
    assert type(params).__name__.endswith("p3_f64")
    cols, n = _cols((rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, q_rim, b_rim, logl), np.float64)
    shift = None if inpc_log_shift is None else np.ascontiguousarray(np.broadcast_to(np.asarray(inpc_log_shift, np.float64), (n,)))
    in13 = (C.c_void_p * 13)(*([_ptr(a) for a in cols] + [_ptr(shift)]))
    out = {k: np.empty(n, np.float64) for k in P3_BMT_OUT}
    o9 = (C.c_void_p * 9)(*[_ptr(out[k]) for k in P3_BMT_OUT])
    st = lib().oracle_bmt2m_p3_f64(C.byref(params), C.c_int64(n), in13, o9, C.c_int(int(bound)))
    assert st == 0
    return out


# ---- 0-moment scheme ---------------------------------------------------------------------------
def bmt0m(params, q_lcl, q_icl, q_vap_sat=None):
    """BMT:658-680 over arrays (native arithmetic of the block's float type)."""
    dtype = np.float64 if type(params).__name__.endswith("f64") else np.float32
    cols = [q_lcl, q_icl] + ([q_vap_sat] if q_vap_sat is not None else [])
    cols, n = _cols(cols, dtype)
    out = np.empty(n, dtype)
    st = getattr(lib(), f"oracle_bmt0m_{_suf(dtype)}")(C.byref(params), C.c_int64(n), _ptr(cols[0]), _ptr(cols[1]),
                                                      _ptr(cols[2]) if len(cols) == 3 else None, _ptr(out))
    assert st == 0
    return out


ICENUC_RATES = {"MohlerDepositionRate": (0, 4), "P3_het_N_i": (1, 4), "INP_concentration_frequency": (2, 2), "het_ice_nucleation": (3, 5)}


def icenuc_rates(params, what, *cols):
    """Multi-argument nucleation rates (same numbering as cumicro_icenuc_rates_*). Returns (out, out2, n_domain_errors)."""
    dtype = np.float64 if type(params).__name__.endswith("f64") else np.float32
    code, nin = ICENUC_RATES[what]
    assert len(cols) == nin
    cols, n = _cols(cols, dtype)
    tbl = (C.c_void_p * 5)(*([_ptr(c) for c in cols] + [None] * (5 - nin)))
    out, out2 = np.empty(n, dtype), np.empty(n, dtype)
    fn = getattr(lib(), f"oracle_icenuc_rates_{_suf(dtype)}")
    fn.restype = C.c_int64
    nerr = fn(C.byref(params), C.c_int(code), C.c_int64(n), tbl, _ptr(out), _ptr(out2))
    return out, out2, int(nerr)


F23_OUT = ("rain_dn_frz", "rain_dq_frz", "cloud_dn_frz", "cloud_dq_frz", "immersion_limit_dn", "deposition_dn", "deposition_dq")


def icenuc_f23(params, rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice, inpc_log_shift=None, bound=False):
    """The F23 / Bigg rates as BMT:998-1075 calls them (same layout as cumicro_icenuc_f23_*)."""
    assert type(params).__name__.endswith("p3_f64")
    cols, n = _cols((rho, T, q_tot, q_lcl, n_lcl, q_rai, n_rai, q_ice, n_ice), np.float64)
    shift = None if inpc_log_shift is None else np.ascontiguousarray(np.broadcast_to(np.asarray(inpc_log_shift, np.float64), (n,)))
    out = {k: np.empty(n, np.float64) for k in F23_OUT}
    st = lib().oracle_icenuc_f23_f64(C.byref(params), C.c_int64(n), (C.c_void_p * 9)(*[_ptr(a) for a in cols]), _ptr(shift),
                                     (C.c_void_p * 7)(*[_ptr(out[k]) for k in F23_OUT]), C.c_int(int(bound)))
    assert st == 0
    return out


ALT_2M = {"acnv_KK2000": 0, "acnv_B1994": 1, "acnv_TC1980": 2, "acnv_LD2004": 3, "accr_KK2000": 4, "accr_B1994": 5, "accr_TC1980": 6}


def alt_2m(params, what, q_lcl, q_rai=None, rho=None, N_d=None, smooth_transition=False):
    """CM2.conv_q_lcl_to_q_rai / accretion of the alternative 2-moment closures (CM2:920-1002)."""
    dtype = np.float64 if type(params).__name__.endswith("f64") else np.float32
    n = np.asarray(q_lcl).shape[0]
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=dtype)
    q_lcl, q_rai, rho, N_d = c(q_lcl), c(q_rai), c(rho), c(N_d)
    out = np.empty(n, dtype)
    st = getattr(lib(), f"oracle_2m_alt_{_suf(dtype)}")(C.byref(params), C.c_int(ALT_2M[what]), C.c_int(int(smooth_transition)), C.c_int64(n),
                                                       _ptr(q_lcl), _ptr(q_rai), _ptr(rho), _ptr(N_d), _ptr(out))
    assert st == 0
    return out


# ---- cloud diagnostics (src/CloudDiagnostics.jl) ---------------------------------------------------------------
def diag_2m(pdf_c, pdf_r, q_lcl, q_rai, N_lcl, N_rai, rho):
    """(radar_reflectivity_2M, effective_radius_2M) over arrays."""
    dtype = np.float64 if type(pdf_c).__name__.endswith("f64") else np.float32
    cols, n = _cols((q_lcl, q_rai, N_lcl, N_rai, rho), dtype)
    Z, reff = np.empty(n, dtype), np.empty(n, dtype)
    st = getattr(lib(), f"oracle_diag_2m_{_suf(dtype)}")(C.byref(pdf_c), C.byref(pdf_r), C.c_int64(n), *[_ptr(a) for a in cols], _ptr(Z), _ptr(reff))
    assert st == 0
    return Z, reff


def diag_1m(params, q_rai, rho):
    dtype = np.float64 if type(params).__name__.endswith("f64") else np.float32
    cols, n = _cols((q_rai, rho), dtype)
    Z = np.empty(n, dtype)
    assert getattr(lib(), f"oracle_diag_1m_{_suf(dtype)}")(C.byref(params), C.c_int64(n), _ptr(cols[0]), _ptr(cols[1]), _ptr(Z)) == 0
    return Z


def diag_reff_lh97(rho_w, rho, q_lcl, N_lcl=None, q_rai=None, N_rai=None, dtype=np.float64):
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=dtype)
    rho, q_lcl, N_lcl, q_rai, N_rai = c(rho), c(q_lcl), c(N_lcl), c(q_rai), c(N_rai)
    out = np.empty(rho.shape[0], dtype)
    crho = C.c_double(rho_w) if dtype == np.float64 else C.c_float(rho_w)
    assert getattr(lib(), f"oracle_diag_reff_lh97_{_suf(dtype)}")(crho, C.c_int64(rho.shape[0]), _ptr(rho), _ptr(q_lcl), _ptr(N_lcl), _ptr(q_rai),
                                                                  _ptr(N_rai), _ptr(out)) == 0
    return out
