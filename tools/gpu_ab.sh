#!/bin/bash
# usage (on the GPU box): tools/gpu_ab.sh <tag> [cases...] -- A/B matrix of step kernels, one JSON line each -> gpurun_out/ab_<tag>.jsonl
# case syntax: label:bench-args (comma separated args)
tag=${1:-ab}; shift
out=gpurun_out/ab_${tag}.jsonl
: > $out
run() {  # label, bench args...
  label=$1; shift
  line=$(python bench.py --no-e2e --no-cpu "$@" 2>gpurun_out/ab_${tag}_${label}.err | tail -1)
  echo "{\"label\": \"$label\", \"line\": $line}" >> $out
  python - "$label" <<PY
import json,sys
try:
    d=json.loads('''$line''')
    r=d["roofline"]
    print(sys.argv[1], "MLUPS %.0f  step %.3f ms  kernel %.3f ms  fold %.3f  frac(step) %.3f" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["other_kernels_ms"]["fold_or_wrap"], r["step_frac_of_roofline"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
if [ $# -eq 0 ]; then
  set -- "fused_256_det:--nx,256,--ny,256,--nz,256,--kbt,0,--steps,20" "fused_256_noise:--nx,256,--ny,256,--nz,256,--steps,20" \
         "fused_512_noise:--steps,10" "fused_512_det:--kbt,0,--steps,10"
fi
for c in "$@"; do
  label=${c%%:*}; args=${c#*:}
  run $label ${args//,/ }
done
