#!/bin/bash
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_demand_probe tools/probes/tma_demand_probe.cu -lcuda && timeout 120 /tmp/tma_demand_probe 384 > gpurun_out/r2_tma_demand_probe.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/stream_probe tools/probes/stream_pattern_probe.cu && timeout 120 /tmp/stream_probe 384 2>&1 | head -3 >> gpurun_out/r2_tma_demand_probe.txt
python -m pytest tests/test_gpu_noise_quality.py tests/test_gpu_parity.py tests/test_droplet_fit.py -x -q -m gpu -s 2>&1 | grep -E "N = |passed|failed|Error" > gpurun_out/r2j_pytest.log
AB_CASES=r1n python tools/ab.py r2j 2 build/libnewA10.so build/libnewM16.so > gpurun_out/r2j_ab.txt 2>&1
cat gpurun_out/r2_tma_demand_probe.txt gpurun_out/r2j_pytest.log gpurun_out/r2j_ab.txt
