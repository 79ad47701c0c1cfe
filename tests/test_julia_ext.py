"""The Julia package extension (ext/CloudMicrophysicsCuMicroExt.jl) cannot be executed in this image (no Julia), so its
layout contract is checked statically: every `C*` mirror struct between the BEGIN-MIRRORS / END-MIRRORS markers must
match the C struct of include/cumicro_params.inc it names — same field names, order, element types, array lengths —
and every `ccall` must name a symbol that include/cumicro.h declares, with as many argument types as the C prototype."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXT = os.path.join(ROOT, "ext", "CloudMicrophysicsCuMicroExt.jl")


def _julia_mirrors():
    src = open(EXT, encoding="utf-8").read()
    body = src[src.index("BEGIN-MIRRORS"):src.index("(END-MIRRORS)")]
    out = {}
    for m in re.finditer(r"#\s*(cumicro_\w+)\s*\nstruct\s+(\w+)(\{FT\})?\s*\n(.*?)\nend", body, flags=re.S):
        cname, jname, _, fields = m.group(1), m.group(2), m.group(3), m.group(4)
        fl = []
        for line in fields.splitlines():
            line = line.strip()
            if not line:
                continue
            fm = re.match(r"(\w+)::(.+)$", line)
            assert fm, (jname, line)
            fl.append((fm.group(1), fm.group(2).strip()))
        out[cname] = (jname, fl)
    return out


def test_every_c_struct_has_a_julia_mirror_with_the_same_layout(built):
    abi = built._abi
    parsed = dict(abi._parse_structs(os.path.join(abi.INCLUDE_DIR, "cumicro_params.inc")))
    mirrors = _julia_mirrors()
    assert set(parsed) == set(mirrors), (sorted(set(parsed) - set(mirrors)), sorted(set(mirrors) - set(parsed)))
    jname_of = {c: j for c, (j, _) in mirrors.items()}
    for cname, cfields in parsed.items():
        jname, jfields = mirrors[cname]
        assert [f for _, f, _ in cfields] == [f for f, _ in jfields], (cname, jname)
        for (tok, fname, alen), (_, jtype) in zip(cfields, jfields):
            if tok == "CUMICRO_FT":
                base = "FT"
            elif tok == "int32_t":
                base = "Int32"
            else:
                nested = jname_of[tok]
                base = nested + ("{FT}" if re.search(r"struct\s+%s\{FT\}" % nested, open(EXT, encoding="utf-8").read()) else "")
            want = base if alen is None else "NTuple{%d, %s}" % (alen, base)
            assert jtype == want, (cname, fname, jtype, want)
        # FT fields before Int32 fields (natural alignment = C layout for both float types)
        kinds = [t for t, _, _ in cfields]
        if "int32_t" in kinds:
            first_int = kinds.index("int32_t")
            assert all(k == "int32_t" for k in kinds[first_int:]), cname


def _c_prototypes(abi):
    src = open(os.path.join(abi.INCLUDE_DIR, "cumicro.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|void|int64_t|const char\*)\s+(cumicro_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        protos[m.group(1)] = len(args)
    return protos


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_every_ccall_names_a_declared_symbol_with_the_right_arity(built):
    abi = built._abi
    protos = _c_prototypes(abi)
    src = open(EXT, encoding="utf-8").read()
    n_calls = 0
    for m in re.finditer(r"ccall\(\((sym\(:(\w+), FT\)|:(\w+)), libcumicro\), (\w+),\s*\((.*?)\)[,)]", src, flags=re.S):
        base = m.group(2) or m.group(3)
        names = [base + "_f64", base + "_f32"] if m.group(2) else [base]
        argtypes = _split_top(m.group(5))
        # `extra_types...` splices (FT, Cint) into the linavg call: count it as the two arguments it stands for
        n_args = sum(2 if a == "extra_types..." else 1 for a in argtypes)
        for nm in names:
            assert nm in protos, f"ccall of undeclared symbol {nm}"
            if "extra_types..." in argtypes:
                assert n_args - 2 <= protos[nm] <= n_args, (nm, n_args, protos[nm])   # inst (no extras) and linavg (dt, nsub) share the call
            else:
                assert n_args == protos[nm], (nm, n_args, protos[nm])
        n_calls += 1
    assert n_calls >= 20, n_calls


def test_array_methods_cover_section_3_4(built):
    """Every stand-alone entry point of SURVEY §3.4 / DESIGN §1 has an array method in the extension."""
    src = open(EXT, encoding="utf-8").read()
    for name in ("BMT.bulk_microphysics_tendencies", "CM2.rain_terminal_velocity", "CM2.cloud_terminal_velocity",
                 "CM2.conv_q_lcl_to_q_rai", "CM2.accretion", "CM1.terminal_velocity", "CMNonEq.terminal_velocity",
                 "CMNonEq.conv_q_vap_to_q_lcl", "CMNonEq.conv_q_vap_to_q_icl", "CM_HetIce.deposition_J", "CM_HetIce.ABIFM_J",
                 "CM_HomIce.homogeneous_J_cubic", "CM_HomIce.homogeneous_J_linear", "CO.a_w_ice", "CO.a_w_eT", "CO.a_w_xT",
                 "CM_HetIce.MohlerDepositionRate", "CM_HetIce.P3_het_N_i", "CM_HetIce.INP_concentration_frequency",
                 "AA.N_activated_per_mode", "AA.M_activated_per_mode", "AA.total_N_activated", "AA.max_supersaturation",
                 "P3.get_distribution_logλ_from_prognostic", "P3.ice_terminal_velocity_number_weighted_from_prognostic",
                 "P3.het_ice_nucleation", "CMD.radar_reflectivity_2M", "CMD.effective_radius_2M", "CMD.radar_reflectivity_1M",
                 "CMD.effective_radius_Liu_Hallet_97", "DomainError", "AssertionError"):
        assert name in src, name
    assert src.count("BMT.bulk_microphysics_tendencies(") >= 6   # 2M warm, 2M+P3, 1M x3, 0M
