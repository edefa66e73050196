#!/usr/bin/env python
"""Dynamic SASS opcode mix + stall samples from an ncu report: tools/ncu_opmix.py <report.ncu-rep> <warp-planes>"""
import collections, csv, subprocess, sys
rep, wp = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not" not in h]
byop, samp, stalls, tot = collections.Counter(), collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) <= ie:
        continue
    try:
        e = int(r[ie])
    except ValueError:
        continue
    t = r[ia].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    byop[op] += e; tot += e; samp[op] += int(r[isamp] or 0)
    for i in stall_cols:
        try:
            stalls[hdr[i]] += int(r[i] or 0)
        except ValueError:
            pass
print(f"total warp instr {tot}  per warp-plane {tot / wp:.1f}")
for op, c in byop.most_common(45):
    print(f"{op:10s} {c / wp:8.1f}  samples {samp[op]}")
print(stalls.most_common(10))
