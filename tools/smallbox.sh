#!/bin/bash
# small-box regime (the reference's own job sizes, Parameters:1-37): MLUPS for brick heights / CTA sizes / graphs on-off
out=gpurun_out/${1:-smallbox}.txt; : > $out
run() {  # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py --no-e2e --no-cpu "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label', '%.1f MLUPS' % d['value'], '%.2f us/step' % (d['ms_per_step']*1e3), 'launches', d['gpu_launches'])" >> $out
}
for lz in 0 2 4 8; do
  run "32^3 det graph lz=$lz" BFLBM_GRAPH=1 -- --nx 32 --ny 32 --nz 32 --kbt 0 --steps 2048 --warmup 128 --brick-lz $lz
  run "8x256x64 noise graph lz=$lz" BFLBM_GRAPH=1 -- --nx 8 --ny 256 --nz 64 --steps 2048 --warmup 128 --brick-lz $lz
done
for lz in 2 4; do
  run "32^3 det graph NT128 lz=$lz" BFLBM_GRAPH=1 BFLBM_CTA_THREADS=128 -- --nx 32 --ny 32 --nz 32 --kbt 0 --steps 2048 --warmup 128 --brick-lz $lz
  run "8x256x64 noise graph NT128 lz=$lz" BFLBM_GRAPH=1 BFLBM_CTA_THREADS=128 -- --nx 8 --ny 256 --nz 64 --steps 2048 --warmup 128 --brick-lz $lz
done
run "32^3 det plain lz=0" BFLBM_GRAPH=0 -- --nx 32 --ny 32 --nz 32 --kbt 0 --steps 2048 --warmup 128
run "8x256x64 noise plain lz=0" BFLBM_GRAPH=0 -- --nx 8 --ny 256 --nz 64 --steps 2048 --warmup 128
run "64^3 noise graph" BFLBM_GRAPH=1 -- --nx 64 --ny 64 --nz 64 --steps 2048 --warmup 128
run "64^3 noise plain" BFLBM_GRAPH=0 -- --nx 64 --ny 64 --nz 64 --steps 2048 --warmup 128
run "128^3 noise graph" BFLBM_GRAPH=1 -- --nx 128 --ny 128 --nz 128 --steps 512 --warmup 64
run "128^3 noise plain" BFLBM_GRAPH=0 -- --nx 128 --ny 128 --nz 128 --steps 512 --warmup 64
cat $out
