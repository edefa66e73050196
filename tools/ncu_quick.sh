#!/bin/bash
# usage (on the GPU box): tools/ncu_quick.sh <tag> <bench args...>   -> gpurun_out/<tag>.csv  (cheap metrics of the fused kernel)
tag=$1; shift
CMD="python bench.py --nx 256 --ny 256 --nz 256 --steps 2 --warmup 1 --no-e2e --no-cpu $@"
$CMD > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio \
  --clock-control none -k regex:k_step_fused -s 2 -c 1 --csv --log-file gpurun_out/${tag}.csv $CMD > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${tag}.csv")) if len(r)>10]
for r in rows[1:]:
    print("${tag}", r[-3], r[-1])
PY
