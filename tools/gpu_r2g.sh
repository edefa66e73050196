#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_droplet_fit.py tests/test_gpu_noise_quality.py tests/test_reference_numbers.py -x -q -m gpu -s 2>&1 | tail -12 > gpurun_out/r2g_pytest.log
AB_CASES=r1n,gn python tools/ab.py r2g 3 build/libnewT0.so build/libnewT1.so > gpurun_out/r2g_ab.txt 2>&1
cat gpurun_out/r2g_pytest.log gpurun_out/r2g_ab.txt
