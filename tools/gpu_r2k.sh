#!/bin/bash
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_demand_probe tools/probes/tma_demand_probe.cu -lcuda
: > gpurun_out/r2_tma_demand_probe.txt
for a in "1 38 0" "1 38 1" "0 38 0"; do set -- $a; timeout 120 /tmp/tma_demand_probe 384 $1 $2 $3 2>&1 | grep -v status >> gpurun_out/r2_tma_demand_probe.txt; done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/stream_probe tools/probes/stream_pattern_probe.cu && timeout 120 /tmp/stream_probe 384 2>&1 | head -1 >> gpurun_out/r2_tma_demand_probe.txt
cat gpurun_out/r2_tma_demand_probe.txt
