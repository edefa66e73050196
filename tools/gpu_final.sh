#!/bin/bash
# usage (GPU box): tools/gpu_final.sh <round-tag>  -- the evidence set of a round: GPU tests, bench lines, launch list, traffic, full captures
tag=${1:-r2}
o=gpurun_out
BFLBM_STATS_OUT=$o python -m pytest tests -m gpu -q 2>&1 | tail -6 > $o/${tag}_pytest_gpu.log; cat $o/${tag}_pytest_gpu.log
python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; tail -c 800 $o/${tag}_bench_n1.json; echo
python bench.py --impl reference > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_ref.err
python bench.py --kbt 0 --no-e2e --no-cpu --steps 20 > $o/${tag}_bench_n1_det.json 2>/dev/null
BFLBM_RATE1=0 python bench.py --no-e2e --no-cpu --steps 20 > $o/${tag}_bench_n1_general.json 2>/dev/null
# the reference's own job sizes (Parameters:1-37), library's choice of kernels, CUDA-graph replay
python bench.py --algo auto --nx 32 --ny 32 --nz 32 --kbt 0 --steps 4096 --warmup 128 --no-e2e --no-cpu > $o/${tag}_bench_32cubed_det.json 2>/dev/null
python bench.py --algo auto --nx 8 --ny 256 --nz 64 --steps 4096 --warmup 128 --no-e2e --no-cpu > $o/${tag}_bench_8x256x64_noise.json 2>/dev/null
python bench.py --algo auto --nx 64 --ny 64 --nz 64 --steps 4096 --warmup 128 --no-e2e --no-cpu > $o/${tag}_bench_64cubed_noise.json 2>/dev/null
python bench.py --algo auto --nx 128 --ny 128 --nz 128 --steps 1024 --warmup 64 --no-e2e --no-cpu > $o/${tag}_bench_128cubed_noise.json 2>/dev/null
# launch list of the bench command (cold-cache, serialised: compare shares)
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > $o/${tag}_launch_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_512cubed.csv $CMD > $o/${tag}_launch_ncu.log 2>&1
# DRAM traffic of one launch of the dominant kernel at the bench size
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > $o/${tag}_traffic_plain.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none -k regex:k_step_fused -s 3 -c 1 --csv --log-file $o/${tag}_traffic_fused_512cubed.csv $CMD > /dev/null 2>&1
# full captures (256^3 keeps the replay short)
for v in noise det general; do
  extra=""; [ $v = det ] && extra="--kbt 0"
  r1=1; [ $v = general ] && r1=0
  CMD="python bench.py --nx 256 --ny 256 --nz 256 --steps 2 --warmup 1 --no-e2e --no-cpu $extra"
  BFLBM_RATE1=$r1 $CMD > $o/${tag}_full_${v}_plain.log 2>&1 && BFLBM_RATE1=$r1 ncu --set full --clock-control none --import-source on -k regex:k_step_fused -s 2 -c 1 -o $o/prof_${tag}_fused_${v} -f $CMD > $o/${tag}_full_${v}_ncu.log 2>&1
done
ls -la $o | tail -20
