// cm_thermo.cuh — moist thermodynamics needed by the tendency kernels.
//
// Device form of what the reference reaches through src/ThermodynamicsInterface.jl
// (TDI:9-33, 60-125), i.e. Thermodynamics.jl's saturation vapour pressure, latent
// heats, cp_m and supersaturation (formulas: SURVEY.md §A.1).
//
// ThermoK holds per-launch constants derived on the host (cm::make_thermo_k) so
// no thread spends FP64 divisions on parameter-only expressions.
#pragma once
#include "cm_types.cuh"

namespace cm {

template <class FT> struct ThermoK {
    FT T_0, T_triple, inv_T_triple, press_triple, T_freeze, R_v, inv_R_v;
    FT cp_d, dcp_vd, dcp_lv, dcp_iv;  // cp_d, cp_v-cp_d, cp_l-cp_v, cp_i-cp_v
    FT LH_v0, LH_s0, dcp_vl, dcp_vi;  // cp_v-cp_l, cp_v-cp_i
    FT a_liq, b_liq, a_ice, b_ice;    // p_sat exponents: dcp/R_v, (LH_0 - dcp T_0)/R_v
    FT cv_l, q_min, LH_f0, dcp_li;    // cp_l; q_min; LH_s0-LH_v0; cp_l-cp_i
    // Numerical thresholds of the METHOD's float type (UT:318-340): eps(FT) and
    // ϵ_numerics(FT) = cbrt(floatmin(FT)).  The Float32 entry points compute in Float64 but
    // gate regimes with the Float32 thresholds, so branch selection is the reference's.
    FT eps, eps_n;
};

template <class FT> __host__ inline ThermoK<FT> make_thermo_k(const typename P<FT>::thermo& t, bool method_is_f32 = false) {
    ThermoK<FT> k;
    k.eps = method_is_f32 ? FT(1.1920928955078125e-07) : FT(2.220446049250313e-16);
    k.eps_n = method_is_f32 ? FT(2.2737367544323206e-13) : FT(2.8126442852362996e-103);
    k.T_0 = t.T_0; k.T_triple = t.T_triple; k.inv_T_triple = FT(1) / t.T_triple;
    k.press_triple = t.press_triple; k.T_freeze = t.T_freeze;
    k.R_v = t.R_v; k.inv_R_v = FT(1) / t.R_v;
    k.cp_d = t.cp_d; k.dcp_vd = t.cp_v - t.cp_d; k.dcp_lv = t.cp_l - t.cp_v; k.dcp_iv = t.cp_i - t.cp_v;
    k.LH_v0 = t.LH_v0; k.LH_s0 = t.LH_s0; k.dcp_vl = t.cp_v - t.cp_l; k.dcp_vi = t.cp_v - t.cp_i;
    k.a_liq = k.dcp_vl / t.R_v; k.b_liq = (t.LH_v0 - k.dcp_vl * t.T_0) / t.R_v;
    k.a_ice = k.dcp_vi / t.R_v; k.b_ice = (t.LH_s0 - k.dcp_vi * t.T_0) / t.R_v;
    k.cv_l = t.cp_l; k.q_min = t.q_min; k.LH_f0 = t.LH_s0 - t.LH_v0; k.dcp_li = t.cp_l - t.cp_i;
    return k;
}

// Temperature-dependent quantities shared by every process at one grid point.
template <class FT> struct TempState {
    FT T, inv_T, log_Tr, dinvT;  // log(T/T_triple), 1/T_triple - 1/T
};

template <class FT> CM_DEV TempState<FT> temp_state(const ThermoK<FT>& k, FT T) {
    TempState<FT> s;
    s.T = T;
    s.inv_T = rcp_(T);
    s.log_Tr = logp_(T * k.inv_T_triple);
    s.dinvT = (T - k.T_triple) * s.inv_T * k.inv_T_triple;  // = 1/T_triple - 1/T without cancellation
    return s;
}

// TD.saturation_vapor_pressure(tps, T, Liquid()/Ice()):
//   p_triple (T/T_triple)^(dcp/R_v) exp((LH_0 - dcp T_0)/R_v (1/T_triple - 1/T))
// evaluated as ONE exponential (one exp_ instead of the reference's pow + exp; the
// exponent is O(10), so the result carries ~1e-15 relative error like the reference's).
template <class FT> CM_DEV FT p_sat_liq(const ThermoK<FT>& k, const TempState<FT>& s) {
    return k.press_triple * exp_(fma_(k.a_liq, s.log_Tr, k.b_liq * s.dinvT));
}
template <class FT> CM_DEV FT p_sat_ice(const ThermoK<FT>& k, const TempState<FT>& s) {
    return k.press_triple * exp_(fma_(k.a_ice, s.log_Tr, k.b_ice * s.dinvT));
}
// The temperature-only part of the thermodynamic state that the 1-moment and the ARG2000 / ice-nucleation bodies both start
// from: computed once per point by the fused kernel (kernels_fused.cu) and handed to both.
template <class FT> struct ThermoShared {
    TempState<FT> ts;
    FT p_vs_l, p_vs_i, inv_pvs_l, inv_pvs_i;   // saturation vapour pressures and 1 / max(p_vs, eps_n)
};
template <class FT> CM_DEV ThermoShared<FT> thermo_shared(const ThermoK<FT>& k, FT T) {
    ThermoShared<FT> s;
    s.ts = temp_state(k, T);
    s.p_vs_l = p_sat_liq(k, s.ts);
    s.p_vs_i = p_sat_ice(k, s.ts);
    s.inv_pvs_l = rcp_(fmax_(s.p_vs_l, k.eps_n));
    s.inv_pvs_i = rcp_(fmax_(s.p_vs_i, k.eps_n));
    return s;
}

template <class FT> CM_DEV FT latent_heat_vapor(const ThermoK<FT>& k, FT T) { return fma_(k.dcp_vl, T - k.T_0, k.LH_v0); }
template <class FT> CM_DEV FT latent_heat_sublim(const ThermoK<FT>& k, FT T) { return fma_(k.dcp_vi, T - k.T_0, k.LH_s0); }
template <class FT> CM_DEV FT latent_heat_fusion(const ThermoK<FT>& k, FT T) { return fma_(k.dcp_li, T - k.T_0, k.LH_f0); }
template <class FT> CM_DEV FT cp_m(const ThermoK<FT>& k, FT qt, FT ql, FT qi) {
    return fma_(k.dcp_iv, qi, fma_(k.dcp_lv, ql, fma_(k.dcp_vd, qt, k.cp_d)));
}
// TDI.q_vap                                                       TDI:60-61
template <class FT> CM_DEV FT q_vap(FT qt, FT ql, FT qi) { return clamp0_(qt - ql - qi); }

// CO.G_func_liquid / G_func_ice                                    CO:47-102
//   1 / (L/K/T (L/R_v/T - 1) + R_v T / D / p_vs); inv_p_vs = 1/max(p_vs, eps) from the caller.
template <class FT>
CM_DEV FT G_func(const ThermoK<FT>& k, FT inv_K_safe, FT inv_D_safe, FT L, FT inv_p_vs, const TempState<FT>& s) {
    const FT LT = L * s.inv_T;
    return rcp_(fma_(LT * inv_K_safe, fma_(LT, k.inv_R_v, FT(-1)), k.R_v * s.T * inv_D_safe * inv_p_vs));
}

}  // namespace cm
