#!/usr/bin/env python3
"""Per-kernel SASS statistics of a built library: instruction count, local-memory (spill)
loads / stores, registers and stack.  usage: sass_stats.py [lib.so] [name filter]"""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "operational-space-control_b200/libosc_b200.so"
flt = sys.argv[2] if len(sys.argv) > 2 else ""
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s*(.*)", res):
    usage[m.group(1)] = " ".join(re.findall(r"(?:REG|STACK|SHARED):\d+", m.group(2)))
name, stats = None, {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1); stats[name] = [0, 0, 0]; continue
    if name and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
        stats[name][0] += 1
        if " LDL" in line: stats[name][1] += 1
        if " STL" in line: stats[name][2] += 1
for k, (n, l, s) in stats.items():
    if flt in k:
        d = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        print(f"{n:7d} instr  LDL {l:3d}  STL {s:3d}  {usage.get(k, '')}  {d[:90]}")
