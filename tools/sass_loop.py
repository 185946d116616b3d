#!/usr/bin/env python3
"""Developer probe: instruction mix of the innermost loops of one kernel (SASS backward
branches).  usage: sass_loop.py lib kernel-substring [max-body]"""
import re, subprocess, sys
lib, flt = sys.argv[1], sys.argv[2]
maxb = int(sys.argv[3]) if len(sys.argv) > 3 else 400
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, ins = None, []
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1); continue
    if name and flt in name:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m: ins.append((int(m.group(1), 16), m.group(2)))
print(f"{len(ins)} instructions")
for a, t in ins:
    m = re.search(r"BRA\S*\s+(?:.*,\s*)?(0x[0-9a-f]+)", t)
    if not m: continue
    tgt = int(m.group(1), 16)
    if tgt >= a: continue
    body = [x for x in ins if tgt <= x[0] <= a]
    if len(body) > maxb or len(body) < 40: continue
    ops = {}
    for _, x in body:
        f = x.split()
        op = (f[1] if f[0].startswith("@") else f[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMMA"))
    print(f"loop {tgt:#x}..{a:#x}: {len(body)} instr, FP64-pipe {fp64}: " +
          " ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])))
