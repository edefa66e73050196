#!/bin/bash
# usage (GPU box): tools/gpu_ab2.sh <tag> <reps> <libA> <libB> [bench args]: interleaved A/B of two builds, median of reps
tag=$1; reps=$2; A=$3; B=$4; shift 4
for r in $(seq $reps); do
  for L in $A $B; do
    BFLBM_LIB=$PWD/$L python bench.py --no-e2e --no-cpu "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$L', '%.3f' % d['ms_per_step'], '%.3f' % d['roofline']['kernel_ms'])" >> gpurun_out/ab2_$tag.txt
  done
done
python - <<PY
import collections,statistics
d=collections.defaultdict(list)
for ln in open('gpurun_out/ab2_$tag.txt'):
    l,ms,k=ln.split(); d[l].append(float(ms))
for l,v in d.items(): print('$tag', l, 'median step ms %.3f' % statistics.median(v), sorted(v))
PY
