#!/usr/bin/env python3
"""Per-kernel SASS statistics of a built library: instruction count, local-memory (spill)
loads / stores, registers and stack, plus the mnemonics that show which units the kernels
use: DMMA (FP64 tensor cores, mma.sync m8n8k4), UBLKCP (1-D TMA bulk copies), DFMA / DMUL /
DADD / DSETP (FP64 pipe), SHFL, LDS / STS, SYNCS (mbarrier), MUFU.
usage: sass_stats.py [lib.so] [name filter]      (committed output: profiles/sass_ops.txt)

tcgen05 / TMEM do not appear and cannot: tcgen05.mma has no FP64 kind (f16 / tf32 / f8f6f4 /
i8 / mxf* only), and every contraction on this path must stay FP64 (FP32 anywhere
in the recursion breaks parity), so the FP64 tensor-core path of sm_100a is mma.sync DMMA."""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "operational-space-control_b200/libosc_b200.so"
flt = sys.argv[2] if len(sys.argv) > 2 else ""
OPS = ("DMMA", "UBLKCP", "DFMA", "DMUL", "DADD", "DSETP", "MUFU", "SHFL", "LDS", "STS", "SYNCS",
       "LDG", "STG", "ATOM", "BAR", "UTMA", "UTCMMA")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s*(.*)", res):
    usage[m.group(1)] = " ".join(re.findall(r"(?:REG|STACK|SHARED):\d+", m.group(2)))
name, stats, ops = None, {}, {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1); stats[name] = [0, 0, 0]; ops[name] = dict.fromkeys(OPS, 0); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line) if name else None
    if m:
        stats[name][0] += 1
        mn = m.group(1)
        if mn.startswith("LDL"): stats[name][1] += 1
        if mn.startswith("STL"): stats[name][2] += 1
        for o in OPS:
            if mn == o or mn.startswith(o + "."):
                ops[name][o] += 1
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print(f"# {lib}: {', '.join(arch)}")
for k, (n, l, s_) in stats.items():
    if flt in k:
        d = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        print(f"{n:7d} instr  LDL {l:3d}  STL {s_:3d}  {usage.get(k, '')}  {d[:100]}")
        print("        " + "  ".join(f"{o} {c}" for o, c in ops[k].items() if c))
