#!/usr/bin/env python
"""Static SASS histogram of the kernels of an object / library whose (mangled) name matches a regex.
usage: sass_hist.py <file.o|.so> <regex> [--dump out.txt]"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], re.compile(sys.argv[2])
dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, kernels = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); kernels[cur] = []; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        kernels[cur].append((m.group(1), m.group(2).strip()))
for name, ins in kernels.items():
    if not pat.search(name):
        continue
    ops = [re.sub(r"^@!?U?P\w+\s+", "", t).split()[0].split(".")[0] for _, t in ins]
    c = collections.Counter(ops)
    fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD"))
    print(name)
    print(f"  total {len(ins)}  FP64 arith {fp64} (DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']})  DSETP {c['DSETP']}  other {len(ins) - fp64 - c['DSETP']}")
    print("  " + ", ".join(f"{k}:{v}" for k, v in c.most_common(30)))
    if dump:
        open(dump, "w").write("\n".join(f"{a} {t}" for a, t in ins) + "\n")
