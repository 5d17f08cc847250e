"""SASS census of libdumux_b200.so: per kernel, the instruction counts that prove which hardware paths the code uses
(UBLKCP = bulk async copy / TMA 1-D, SYNCS = mbarrier operations, DFMA / DMUL / DADD = the FP64 mix under -fmad=false,
LDG/STG/LDS/STS, local-memory spills LDL/STL, BAR).    usage: python scripts/sass_census.py [out=profiles/sass_census.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_census.txt")
so = os.path.join(ROOT, "dumux_b200", "libdumux_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
OPS = ["UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR", "ATOM", "RED", "SHFL"]
per = collections.OrderedDict()
cur = None
arch = None
for line in txt.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void dmx::", "").replace("dmx::", "")
        cur = per.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        cur["total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                cur[o] += 1
        if ".TRANS64" in line or "ARRIVE.TRANS" in line:
            cur["SYNCS.TRANS64"] += 1
with open(out, "w") as f:
    f.write(f"# cuobjdump -sass dumux_b200/libdumux_b200.so ({arch}); static instruction counts per kernel\n")
    cols = ["total"] + OPS
    f.write("kernel".ljust(64) + "".join(c.rjust(8) for c in cols) + "\n")
    for name, c in per.items():
        f.write(name[:63].ljust(64) + "".join(str(c.get(k, 0)).rjust(8) for k in cols) + "\n")
print(open(out).read()[:6000])
