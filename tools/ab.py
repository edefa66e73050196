#!/usr/bin/env python
"""Interleaved A/B of library builds on one GPU box (box-to-box variation is +-3 %, inside one call runs agree to 0.1 %).
usage: tools/ab.py <tag> <reps> lib1.so lib2.so ... [-- bench args]
Cases: rate-1 / general-rate (BFLBM_RATE1=0) x kBT = 1e-5 / 0.  Writes gpurun_out/ab_<tag>.json and prints medians."""
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
argv = sys.argv[1:]
extra = []
if "--" in argv:
    i = argv.index("--")
    argv, extra = argv[:i], argv[i + 1:]
tag, reps, libs = argv[0], int(argv[1]), argv[2:]
cases = os.environ.get("AB_CASES", "r1n,r1d,gn,gd").split(",")
CASE = {"r1n": ({}, ["--kbt", "1e-5"]), "r1d": ({}, ["--kbt", "0"]), "gn": ({"BFLBM_RATE1": "0"}, ["--kbt", "1e-5"]),
        "gd": ({"BFLBM_RATE1": "0"}, ["--kbt", "0"])}
res = {}
for r in range(reps):
    for c in cases:
        for lib in libs:
            env = dict(os.environ, BFLBM_LIB=os.path.join(ROOT, lib), **CASE[c][0])
            p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-e2e", "--no-cpu", "--steps", "10", "--warmup", "3"] + CASE[c][1] + extra,
                               capture_output=True, text=True, env=env)
            try:
                d = json.loads(p.stdout.strip().splitlines()[-1])
                res.setdefault(c, {}).setdefault(lib, []).append((d["ms_per_step"], d["roofline"]["kernel_ms"], d["clocks"]["sm_mhz"]))
            except Exception as e:  # noqa
                res.setdefault(c, {}).setdefault(lib, []).append(("fail", p.stderr[-300:]))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"ab_{tag}.json"), "w"), indent=1)
for c, byl in res.items():
    for lib, v in byl.items():
        ok = [x for x in v if x[0] != "fail"]
        if ok:
            print(f"{tag} {c:4s} {lib:28s} step {statistics.median(x[0] for x in ok):7.3f} ms  kernel {statistics.median(x[1] for x in ok):7.3f} ms  "
                  f"clk {statistics.median(x[2] for x in ok if x[2]) if any(x[2] for x in ok) else 0:.0f}  n={len(ok)}")
        else:
            print(tag, c, lib, "FAILED", v[:1])
