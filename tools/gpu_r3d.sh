#!/bin/bash
# sustained 200-step runs: plain launches with per-kernel events (the bench's timed region) against graph replays (what bflbm_step does
# for a caller), and 20 steps for comparison -- where do the 17.1 ms/step of the e2e leg come from?
o=gpurun_out
python bench.py --steps 20 --no-e2e --no-cpu > $o/r3d_steps20_events.json 2>/dev/null
python bench.py --steps 200 --no-e2e --no-cpu > $o/r3d_steps200_events.json 2>/dev/null
BFLBM_BENCH_PROFILE_IN_TIMED=0 python bench.py --steps 200 --no-e2e --no-cpu > $o/r3d_steps200_graph.json 2>/dev/null
BFLBM_BENCH_PROFILE_IN_TIMED=0 BFLBM_GRAPH=0 python bench.py --steps 200 --no-e2e --no-cpu > $o/r3d_steps200_plain.json 2>/dev/null
BFLBM_BENCH_PROFILE_IN_TIMED=0 python bench.py --steps 20 --no-e2e --no-cpu > $o/r3d_steps20_graph.json 2>/dev/null
for f in steps20_events steps200_events steps200_graph steps200_plain steps20_graph; do python - <<PY
import json
d=json.load(open("$o/r3d_$f.json"))
print("$f", round(d["ms_per_step"],3), d["clocks"], d["roofline"]["kernel_ms"] if d["roofline"] else None)
PY
done
