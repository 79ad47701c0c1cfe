#!/usr/bin/env python
"""Executed-instruction shares per source line and per opcode of one captured kernel (ncu --import-source on, -lineinfo):
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [n_points] [top]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
npts = float(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = hdr = curline = None
agg, ops, src = collections.Counter(), collections.Counter(), {}
tot = 0


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 3 and r[0] == "Line No":
        hdr = r
        iex = hdr.index("Instructions Executed")
    elif hdr is not None and len(r) > iex:
        if r[0].isdigit():
            curline = (cur, int(r[0]))
            src[curline] = r[1]
        elif r[0] == "":
            e = I(r[iex])
            m = re.match(r"(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[3].strip())
            if m:
                ops[m.group(1).split(".")[0]] += e
            agg[curline] += e
            tot += e
print(f"executed warp instructions {tot}" + (f" = {tot / (npts / 32):.1f} per 32 points" if npts else ""))
print("  ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(16)))
for (f, l), e in agg.most_common(top):
    print(f"{100 * e / tot:5.1f}%  {f}:{l}  {src[(f, l)].strip()[:110]}")
