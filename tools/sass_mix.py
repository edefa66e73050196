#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel of libbflbm.so:  tools/sass_mix.py <substring of mangled name> [--dump]"""
import collections
import os
import re
import subprocess
import sys

lib = os.environ.get("BFLBM_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "binary-fluctuating-lattice-boltzmann_b200", "libbflbm.so")
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist, lines = None, collections.Counter(), []
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m:
            ins = m.group(1)
            lines.append(ins)
            t = ins.split()
            op = t[1] if t[0].startswith("@") else t[0]
            hist[op.split(".")[0]] += 1
print(sum(hist.values()), "instructions")
for op, c in hist.most_common(80):
    print(f"{op:10s} {c}")
if "--dump" in sys.argv:
    print("\n".join(lines))
