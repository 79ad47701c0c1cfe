#!/usr/bin/env python
"""Executed-instruction shares per SASS function body (the kernel and each out-of-line device function it calls) of one captured
kernel: python tools/ncu_funcs.py gpurun_out/prof.ncu-rep.  Bodies are split at RET / EXIT; each is labelled by its hottest lines."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = hdr = curline = None
inst = {}   # address -> (sass, executed, thread-executed, (file, line), source text, stall samples)
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 3 and r[0] == "Line No":
        hdr = r
        iex, ith, ism = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    elif hdr is not None and len(r) > iex:
        if r[0].isdigit():
            curline = (cur, int(r[0]), r[1].strip())
        elif r[0] == "" and r[2].startswith("0x"):
            inst[int(r[2], 16)] = (r[3].strip(), int(r[iex] or 0), int(r[ith] or 0), curline, int(r[ism] or 0))
addrs = sorted(inst)
segs, cur_seg = [], []
for a in addrs:
    cur_seg.append(a)
    op = inst[a][0].split()[0] if not inst[a][0].startswith("@") else inst[a][0].split()[1]
    if op.startswith("RET") or op == "EXIT":
        nxt = addrs[addrs.index(a) + 1] if a != addrs[-1] else None
        # a body ends at its LAST RET/EXIT: keep going while later instructions still branch back (cheap test: next is not a body start)
        segs.append(cur_seg)
        cur_seg = []
if cur_seg:
    segs.append(cur_seg)
# merge tiny tails (bodies with several RETs) into the previous body when they share source lines with it
merged = []
for s in segs:
    lines = {inst[a][3][:2] for a in s if inst[a][3]}
    if merged and lines & merged[-1][1] and len(s) < 400:
        merged[-1][0].extend(s)
        merged[-1][1] |= lines
    else:
        merged.append([s, lines])
tot = sum(v[1] for v in inst.values()) or 1
tots = sum(v[4] for v in inst.values()) or 1
print(f"executed warp instructions {tot}, stall samples {tots}")
for s, _ in sorted(merged, key=lambda m: -sum(inst[a][1] for a in m[0])):
    e = sum(inst[a][1] for a in s)
    sm = sum(inst[a][4] for a in s)
    if e < 0.003 * tot:
        continue
    th = sum(inst[a][2] for a in s)
    c = collections.Counter()
    for a in s:
        if inst[a][3]:
            c[inst[a][3]] += inst[a][1]
    top = "; ".join(f"{f}:{l} {t[:50]}" for (f, l, t), _ in c.most_common(3))
    print(f"{100 * e / tot:5.1f}% inst {100 * sm / tots:5.1f}% samples  {len(s):5d} SASS  {th / max(e, 1):4.1f} thr/inst  0x{s[0] & 0xfffff:05x}  {top}")

# ---- the largest body by address chunks: where in the kernel the issue slots and the stall samples go
if len(sys.argv) > 2:
    chunk = int(sys.argv[2])
    body = max(merged, key=lambda m: sum(inst[a][1] for a in m[0]))[0]   # the body that executes the most (the kernel itself)
    print(f"\nthe busiest body in chunks of {chunk} SASS instructions: share of all executed instructions | of all stall samples | cm_p3.cuh / kernels line span")
    for i in range(0, len(body), chunk):
        s = body[i:i + chunk]
        e = sum(inst[a][1] for a in s)
        sm = sum(inst[a][4] for a in s)
        ls = sorted(inst[a][3][1] for a in s if inst[a][3] and inst[a][3][0] in ("cm_p3.cuh",))
        ks = sorted(inst[a][3][1] for a in s if inst[a][3] and inst[a][3][0].startswith("kernels_"))
        span = f"cm_p3.cuh:{ls[len(ls) // 10]}-{ls[-1 - len(ls) // 10]}" if ls else ""
        span += f" kernels:{ks[0]}-{ks[-1]}" if ks else ""
        print(f"  0x{s[0] & 0xfffff:05x}  {100 * e / tot:5.1f}%  {100 * sm / tots:5.1f}%  {span}")
