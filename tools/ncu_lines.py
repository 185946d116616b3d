#!/usr/bin/env python3
"""Per-source-line summary of an ncu report (developer tool).
usage: ncu_lines.py report.ncu-rep [file-substring] [min-percent]"""
import csv, subprocess, sys, collections
def I(x):
    try: return int(x)
    except ValueError: return 0
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else "osc_core3"; minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs = []; cur = None; i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        cur = {"file": r[1], "func": rows[i+1][1], "hdr": rows[i+2], "rows": []}; secs.append(cur); i += 3; continue
    if cur is not None and r: cur["rows"].append(r)
    i += 1
tot_s = tot_i = 0
for s in secs:
    ix = {}
    for k, n in enumerate(s["hdr"]): ix.setdefault(n, k)
    s["ix"] = ix
    for r in s["rows"]:
        if r[0] != "":
            tot_s += I(r[ix["# Samples"]]); tot_i += I(r[ix["Instructions Executed"]])
print("total samples", tot_s, "warp instr", tot_i)
op = collections.Counter()
for s in secs:
    ix = s["ix"]
    for r in s["rows"]:
        if r[0] == "" and len(r) > 3:
            src = r[3].strip()
            if src.startswith("@"): src = src.split(None, 1)[1] if " " in src else src
            o = src.split()[0].split(".")[0] if src else "?"
            op[o] += I(r[ix["Instructions Executed"]])
print("opcodes:", ", ".join(f"{o} {100*c/tot_i:.1f}%" for o, c in op.most_common(16)))
for s in secs:
    if want not in s["file"]: continue
    ix = s["ix"]
    print("==", s["file"], s["func"][:60])
    for r in s["rows"]:
        if r[0] == "": continue
        sm = I(r[ix["# Samples"]]); ins = I(r[ix["Instructions Executed"]])
        if 100*sm/tot_s >= minp:
            print(f"{I(r[0]):5d} {100*sm/tot_s:5.2f}% ins {100*ins/tot_i:5.2f}% sb {I(r[ix['stall_short_sb']]):5d} wait {I(r[ix['stall_wait']]):5d} | {r[1][:90]}")
# region buckets (function-level) for osc_core3.cuh
import re
src = open(sys.argv[4]).read().splitlines() if len(sys.argv) > 4 else None
if src:
    starts = []
    for n, line in enumerate(src, 1):
        m = re.match(r"\s*static OSC_HD .*? (\w+)\(", line)
        if m: starts.append((n, m.group(1)))
    def region(ln):
        name = "?"
        for n, f in starts:
            if n <= ln: name = f
        return name
    agg = collections.Counter(); aggi = collections.Counter()
    for s in secs:
        if want not in s["file"]: continue
        ix = s["ix"]
        for r in s["rows"]:
            if r[0] == "": continue
            agg[region(int(r[0]))] += I(r[ix["# Samples"]]); aggi[region(int(r[0]))] += I(r[ix["Instructions Executed"]])
    print("regions (samples%, instr%):")
    for k, v in agg.most_common(): print(f"  {k:28s} {100*v/tot_s:5.1f}% {100*aggi[k]/tot_i:5.1f}%")
    other = collections.Counter()
    for s in secs:
        if want in s["file"]: continue
        ix = s["ix"]
        for r in s["rows"]:
            if r[0] != "": other[s["file"].split("/")[-1]] += I(r[ix["# Samples"]])
    print("other files:", {k: round(100*v/tot_s, 1) for k, v in other.items()})
