#!/usr/bin/env python3
"""Print the SASS of one kernel between two addresses.  usage: sass_range.py lib filter lo hi"""
import re, subprocess, sys
lib, flt, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1); continue
    if name and flt in name:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and lo <= int(m.group(1), 16) <= hi:
            print(m.group(1), m.group(2))
