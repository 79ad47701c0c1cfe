#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel opcode histogram (static) of the hot kernels of libcumicro.so plus the full
listing of the headline 2M kernel (the SHIPPED instantiation: the regexes name it exactly).  Run after a build:
    python tools/sass_summary.py [round-tag]  ->  profiles/<tag>_sass_*.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "cloudmicrophysics.jl_b200", "build")
KERNELS = [  # (object, regex on the mangled name, label)
    # the launch bmt2m_warm_impl makes for a default-structure limited-PSD block with all four outputs: BLOCK 128, MINB 6, ALL_OUT, PPT 1, TAB
    ("kernels_2m.o", r"warm2m_tile_kernelIdLi7ELi1ELi128ELi6ELb1ELi1ELb1E", "2m_warm_f64"),
    ("kernels_2m.o", r"warm2m_tile_kernelIfLi7ELi1ELi128ELi6ELb1ELi1ELb1E", "2m_warm_f32"),
    # 1M Instantaneous: tile shape, default exponent structure (STD), 128x7, ALL_OUT
    ("kernels_1m.o", r"pointwise_kernel_tiledIdLi7ELi4E.*OneMInstILb1EEELi128ELi7ELb1E", "1m_inst_f64"),
    ("kernels_1m.o", r"pointwise_kernelIdLi7ELi4E.*OneMLinAvgILb1EEELb0", "1m_linavg_f64"),
    # config 3: tile shape, 3 modes, no M_act, 896x1, ALL_OUT
    ("kernels_icenuc.o", r"pointwise_kernel_tiledIfLi8ELi11E.*ArgIceNucILi3ELb0EEELi896ELi1ELb1E", "arg_icenuc_f32"),
    # config 5: 896x1, SPEC 1, TAB, S1M, ALL_OUT, NM3
    ("kernels_fused.o", r"fused_kernelIdLi896ELi1ELb0ELi1ELb1ELb1ELb1ELb1E", "fused_f64"),
    ("kernels_p3.o", r"p3_tile_kernelIdLi0", "p3_rates_f64"),
    ("kernels_emulator.o", r"emu_kernelId", "emulator_f64"),
]
FP64 = ("DFMA", "DMUL", "DADD")


def functions(obj):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True, check=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur is not None:
            out[cur].append(line)
    return out


def opcodes(lines):
    h = collections.Counter()
    for l in lines:
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", l)
        if m:
            h[m.group(1)] += 1
    return h


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    summary = ["static SASS opcode counts of the hot kernels (sm_100a, nvcc 12.9, -fmad=false); dynamic counts are in the ncu summaries",
               "the tile kernels fetch their input tiles with bulk asynchronous copies (UBLKCP.S.G + SYNCS mbarrier waits, one ELECT-ed lane per",
               "issuer warp); the emulator's dense layers are DMMA.8x8x4 (FP64 tensor cores); tcgen05 / tensor-map TMA mnemonics are absent by",
               "design: the tendency kernels are pointwise, not contractions, and their tiles are 1-D runs of a column", ""]
    cache = {}
    for obj, pat, label in KERNELS:
        fns = cache.setdefault(obj, functions(obj))
        names = [n for n in fns if re.search(pat, n)]
        if not names:
            summary.append(f"{label}: kernel not found ({pat})")
            continue
        assert len(names) == 1, (label, names)   # exactly the shipped instantiation, never "the first match"
        lines = fns[names[0]]
        h = opcodes(lines)
        total = sum(h.values())
        fp64 = sum(h[k] for k in FP64)
        summary.append(f"{label}: {total} instructions, {fp64} FP64 arithmetic (DFMA {h['DFMA']}, DMUL {h['DMUL']}, DADD {h['DADD']}), "
                       f"DSETP {h['DSETP']}, MUFU {h['MUFU']}, LDS {h['LDS']}, LDG {h['LDG']}, LDGSTS {h['LDGSTS']}, STG {h['STG']}, "
                       f"LDL {h['LDL']}, STL {h['STL']}, BRA {h['BRA']}, CALL {h['CALL']}, UBLKCP {h['UBLKCP']}, DMMA {h['DMMA']}")
        summary.append("    " + "  ".join(f"{k} {v}" for k, v in h.most_common(24)))
        summary.append(f"    {names[0][:150]}")
        if label == "2m_warm_f64":
            listing = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in lines if not re.match(r"^\s*/\* 0x[0-9a-f]+ \*/\s*$", l)]
            open(os.path.join(ROOT, "profiles", f"{tag}_sass_2m_warm_f64.txt"), "w").write("\n".join(listing) + "\n")
    open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt"), "w").write("\n".join(summary) + "\n")
    print("\n".join(summary))


if __name__ == "__main__":
    sys.exit(main())
