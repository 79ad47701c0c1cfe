#!/usr/bin/env python
"""Static SASS census of one kernel in libcumicro.so: instruction counts by class.
usage: sass_census.py <regex on the (mangled) kernel name> [points per loop body]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "cloudmicrophysics.jl_b200", "libcumicro.so")
pat = re.compile(sys.argv[1])
ppb = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, kernels = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        kernels[cur].append(m.group(1))
for name, ins in kernels.items():
    if not pat.search(name):
        continue
    c = collections.Counter(i.split(".")[0] for i in ins)
    fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    mufu = c.get("MUFU", 0)
    print(f"{name}\n  total {len(ins)}  FP64-pipe {fp64} (DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']} DSETP {c['DSETP']} DMNMX {c.get('DMNMX',0)})"
          f"  MUFU {mufu}  LDG {c.get('LDG',0)} STG {c.get('STG',0)} LDS {c.get('LDS',0)} BRA {c.get('BRA',0)} CALL {c.get('CALL',0)}")
    print(f"  per point (/{ppb}): FP64 {fp64/ppb:.0f}  flops {(2*c['DFMA']+c['DMUL']+c['DADD'])/ppb:.0f}  all {len(ins)/ppb:.0f}")
    print("  top:", ", ".join(f"{k}:{v}" for k, v in c.most_common(14)))
