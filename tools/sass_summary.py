#!/usr/bin/env python
"""Per-kernel summary of libhm_b200/libhmrecon.so (sm_100a): registers, shared memory, instruction count and an opcode histogram
with the Blackwell-relevant mnemonics called out (UBLKCP = TMA bulk copy, LDGSTS = cp.async, UTMALDG / UTC*MMA = tensor TMA / tcgen05
— absent by design: the path is integer stencil work), from `cuobjdump -sass` and `cuobjdump -res-usage`.
usage: tools/sass_summary.py [libhmrecon.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "libhm_b200", "libhmrecon.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2)))
kern, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        kern[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        kern[name][m.group(1)] += 1
KEY = ["UBLKCP", "UTMALDG", "UTMASTG", "UTCHMMA", "UTCIMMA", "UTCQMMA", "LDGSTS", "SYNCS", "IDP", "IMAD", "VIMNMX", "VIADDMNMX", "LDS", "STS", "LDG", "STG", "LDL", "STL",
       "SHFL", "REDUX", "BAR", "MEMBAR", "ATOM", "ATOMG", "NANOSLEEP", "CCTL", "DEPBAR", "LDGDEPBAR", "BRA"]
demangle = subprocess.run(["c++filt"] + list(kern), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS summary of {os.path.basename(so)} (cuobjdump -sass / -res-usage; tools/sass_summary.py)")
for (n, c), d in zip(kern.items(), demangle):
    reg, sh = usage.get(n, (0, 0))
    tot = sum(c.values())
    print(f"\n{d.split('(')[0]}\n  registers {reg}  static smem {sh} B  instructions {tot}")
    print("  " + "  ".join(f"{k}:{c[k]}" for k in KEY if c[k]))
    rest = [(k, v) for k, v in c.most_common(12)]
    print("  top: " + " ".join(f"{k}={v}" for k, v in rest))
