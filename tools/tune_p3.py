#!/usr/bin/env python
"""Launch-shape sweep of the P3 kernel.  `build` (here, no GPU): compile kernels_p3.cu once per variant and link each against the
other objects into cloudmicrophysics.jl_b200/build/variants/libcumicro_<tag>.so.  `run` (on the GPU box): time process rates and
velocities with every variant (CUMICRO_LIB) in a fresh process each."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR = os.path.join(ROOT, "cloudmicrophysics.jl_b200", "build", "variants")
VARIANTS = {
    "b512x2_e1": "-DCUMICRO_P3_BLOCK=512 -DCUMICRO_P3_MINB=2 -DCUMICRO_P3_SYNC_EVERY=1",
    "b512x2_e2": "-DCUMICRO_P3_BLOCK=512 -DCUMICRO_P3_MINB=2 -DCUMICRO_P3_SYNC_EVERY=2",
    "b512x2_e4": "-DCUMICRO_P3_BLOCK=512 -DCUMICRO_P3_MINB=2 -DCUMICRO_P3_SYNC_EVERY=4",
    "b512x2_e999": "-DCUMICRO_P3_BLOCK=512 -DCUMICRO_P3_MINB=2 -DCUMICRO_P3_SYNC_EVERY=999",
    "b256x4_e1": "-DCUMICRO_P3_BLOCK=256 -DCUMICRO_P3_MINB=4 -DCUMICRO_P3_SYNC_EVERY=1",
    "b256x4_e2": "-DCUMICRO_P3_BLOCK=256 -DCUMICRO_P3_MINB=4 -DCUMICRO_P3_SYNC_EVERY=2",
    "b320x3_e1": "-DCUMICRO_P3_BLOCK=320 -DCUMICRO_P3_MINB=3 -DCUMICRO_P3_SYNC_EVERY=1",
    "b1024x1_e1": "-DCUMICRO_P3_BLOCK=1024 -DCUMICRO_P3_MINB=1 -DCUMICRO_P3_SYNC_EVERY=1",
    "b1024x1_s2": "-DCUMICRO_P3_BLOCK=1024 -DCUMICRO_P3_MINB=1 -DCUMICRO_P3_SYNC=2",
    "b512x2_s2": "-DCUMICRO_P3_BLOCK=512 -DCUMICRO_P3_MINB=2 -DCUMICRO_P3_SYNC=2",
    "b768x1_e1": "-DCUMICRO_P3_BLOCK=768 -DCUMICRO_P3_MINB=1 -DCUMICRO_P3_SYNC_EVERY=1",
    "b896x1_e1": "-DCUMICRO_P3_BLOCK=896 -DCUMICRO_P3_MINB=1 -DCUMICRO_P3_SYNC_EVERY=1",
    "series4": "-DP3_SERIES_TEST_EVERY=4",
    "divm": "-DP3_DIV_MARKSTEIN=1",
    "base": "",
}
if len(sys.argv) > 2:
    VARIANTS = {k: v for k, v in VARIANTS.items() if k in sys.argv[2:]}

if sys.argv[1] == "build":
    import __graft_entry__ as g
    g.build()
    os.makedirs(VAR, exist_ok=True)
    others = [os.path.join(g.OBJ, f) for f in os.listdir(g.OBJ) if f.endswith(".o") and f != "kernels_p3.o"]

    def one(item):
        tag, flags = item
        obj = os.path.join(VAR, f"kernels_p3_{tag}.o")
        r = subprocess.run([g._nvcc()] + g.NVCC_FLAGS + flags.split() + ["-c", os.path.join(g.CSRC, "kernels_p3.cu"), "-o", obj], capture_output=True, text=True)
        if r.returncode:
            return tag, r.stderr[-2000:]
        lib = os.path.join(VAR, f"libcumicro_{tag}.so")
        r = subprocess.run([g._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, obj] + others + ["-lcudart"], capture_output=True, text=True)
        os.remove(obj)
        return tag, r.stderr[-2000:] if r.returncode else "ok"
    with ThreadPoolExecutor(8) as ex:
        for tag, msg in ex.map(one, VARIANTS.items()):
            print(tag, msg)
else:
    log2n = sys.argv[1] if sys.argv[1].isdigit() else "20"
    for tag in VARIANTS:
        lib = os.path.join(VAR, f"libcumicro_{tag}.so")
        if not os.path.exists(lib):
            continue
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "p3_profile_driver.py"), "20"], env=dict(os.environ, CUMICRO_LIB=lib),
                           capture_output=True, text=True)
        print(tag, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:], flush=True)
