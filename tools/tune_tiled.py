#!/usr/bin/env python
"""A/B of the launch shapes of the streaming families: `build` (here) compiles kernels_1m.cu / kernels_icenuc.cu / kernels_fused.cu
per variant into cloudmicrophysics.jl_b200/build/variants/libcumicro_<tag>.so; `run` (GPU box) times tools/bench_families.py's
streaming lines with each."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR = os.path.join(ROOT, "cloudmicrophysics.jl_b200", "build", "variants")
VARIANTS = {
    "base": "",
    "1m8_arg128x6": "-DCUMICRO_1M_MINB=8 -DCUMICRO_ARG_BLOCK=128 -DCUMICRO_ARG_MINB=6",
    "1m6_arg128x7": "-DCUMICRO_1M_MINB=6 -DCUMICRO_ARG_BLOCK=128 -DCUMICRO_ARG_MINB=7",
    "1m256x4_arg256x3": "-DCUMICRO_1M_BLOCK=256 -DCUMICRO_1M_MINB=4 -DCUMICRO_1MV_MINB=4 -DCUMICRO_ARG_BLOCK=256 -DCUMICRO_ARG_MINB=3",
    "arg1024x1": "-DCUMICRO_ARG_BLOCK=1024 -DCUMICRO_ARG_MINB=1",
}
FILES = ("kernels_1m.cu", "kernels_icenuc.cu", "kernels_fused.cu")
if len(sys.argv) > 2:
    VARIANTS = {k: v for k, v in VARIANTS.items() if k in sys.argv[2:]}

if sys.argv[1] == "build":
    import __graft_entry__ as g
    g.build()
    os.makedirs(VAR, exist_ok=True)
    others = [os.path.join(g.OBJ, f) for f in os.listdir(g.OBJ) if f.endswith(".o") and f[:-2] + ".cu" not in FILES]

    def one(item):
        tag, flags = item
        objs = []
        for f in FILES:
            obj = os.path.join(VAR, f"{f[:-3]}_{tag}.o")
            r = subprocess.run([g._nvcc()] + g.NVCC_FLAGS + flags.split() + ["-c", os.path.join(g.CSRC, f), "-o", obj], capture_output=True, text=True)
            if r.returncode:
                return tag, r.stderr[-3000:]
            objs.append(obj)
        lib = os.path.join(VAR, f"libcumicro_{tag}.so")
        r = subprocess.run([g._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + others + ["-lcudart"], capture_output=True, text=True)
        for o in objs:
            os.remove(o)
        return tag, r.stderr[-2000:] if r.returncode else "ok"
    with ThreadPoolExecutor(4) as ex:
        for tag, msg in ex.map(one, VARIANTS.items()):
            print(tag, msg)
else:
    import json
    for tag in VARIANTS:
        lib = os.path.join(VAR, f"libcumicro_{tag}.so")
        if not os.path.exists(lib):
            continue
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_families.py")], env=dict(os.environ, CUMICRO_LIB=lib, CUMICRO_FAMILIES_SKIP_P3="1"),
                           capture_output=True, text=True)
        row = []
        for line in r.stdout.splitlines():
            if line.startswith("{"):
                d = json.loads(line)
                if any(s in d["family"] for s in ("1M Inst", "config 3", "config 5", "LinearizedAverage", "Verbose")):
                    row.append(f"{d['family'][:28]}: {d['ms']:.4f}")
        print(tag, " | ".join(row) if row else r.stderr[-800:], flush=True)
