#!/usr/bin/env python
"""Launch-shape A/B of the secondary kernels (tools/bench_secondary.py): build variants here, run them on the GPU box."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR = os.path.join(ROOT, "cloudmicrophysics.jl_b200", "build", "variants")
VARIANTS = {
    "p3l_1024": "-DCUMICRO_P3L_BLOCK=1024 -DCUMICRO_P3L_MINB=1",
    "p3l_896": "-DCUMICRO_P3L_BLOCK=896 -DCUMICRO_P3L_MINB=1",
    "p3l_768": "-DCUMICRO_P3L_BLOCK=768 -DCUMICRO_P3L_MINB=1",
    "p3l_512x2": "-DCUMICRO_P3L_BLOCK=512 -DCUMICRO_P3L_MINB=2",
    "p3l_640": "-DCUMICRO_P3L_BLOCK=640 -DCUMICRO_P3L_MINB=1",
}
FILES = ("kernels_2m.cu", "kernels_p3.cu")

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
    with ThreadPoolExecutor(3) as ex:
        for tag, msg in ex.map(one, VARIANTS.items()):
            print(tag, msg)
else:
    for tag in VARIANTS:
        lib = os.path.join(VAR, f"libcumicro_{tag}.so")
        if not os.path.exists(lib):
            continue
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_secondary.py")], env=dict(os.environ, CUMICRO_LIB=lib), capture_output=True, text=True)
        print(tag, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-600:], flush=True)
