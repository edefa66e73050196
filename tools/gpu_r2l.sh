#!/bin/bash
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_mini tools/probes/tma_mini_probe.cu -lcuda
for m in 0 1 2 3 4 5; do timeout 60 /tmp/tma_mini $m 2>&1 | tail -1; done | tee gpurun_out/r2l_tma_mini.txt
