#!/bin/bash
# small boxes: fused brick kernel vs thread-per-cell two-pass kernels, CUDA graphs on (the cross-over decides the automatic choice)
out=gpurun_out/${1:-smallbox2}.txt; : > $out
run() {
  label=$1; shift
  python bench.py --no-e2e --no-cpu "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label', '%.1f MLUPS' % d['value'], '%.2f us/step' % (d['ms_per_step']*1e3), 'launches', d['gpu_launches'])" >> $out
}
for algo in fused twopass; do
  run "32^3 det $algo" --algo $algo --nx 32 --ny 32 --nz 32 --kbt 0 --steps 4096 --warmup 128
  run "32^3 noise $algo" --algo $algo --nx 32 --ny 32 --nz 32 --steps 4096 --warmup 128
  run "8x256x64 noise $algo" --algo $algo --nx 8 --ny 256 --nz 64 --steps 4096 --warmup 128
  run "64^3 noise $algo" --algo $algo --nx 64 --ny 64 --nz 64 --steps 2048 --warmup 128
  run "96^3 noise $algo" --algo $algo --nx 96 --ny 96 --nz 96 --steps 1024 --warmup 64
  run "128^3 noise $algo" --algo $algo --nx 128 --ny 128 --nz 128 --steps 512 --warmup 64
  run "256^3 noise $algo" --algo $algo --nx 256 --ny 256 --nz 256 --steps 64 --warmup 8
done
BFLBM_GRAPH=0 run "32^3 det twopass plain" --algo twopass --nx 32 --ny 32 --nz 32 --kbt 0 --steps 4096 --warmup 128
BFLBM_GRAPH=0 run "8x256x64 noise twopass plain" --algo twopass --nx 8 --ny 256 --nz 64 --steps 4096 --warmup 128
cat $out
