#!/bin/bash
for L in A16 A10; do
  BFLBM_LIB=$PWD/build/libnew$L.so python -m pytest tests/test_gpu_noise_quality.py -x -q -m gpu -s 2>&1 | grep -E "N = |chi2|passed|failed" | sed "s/^/$L: /"
done > gpurun_out/r2h_noise.log
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2 >> gpurun_out/r2h_noise.log
AB_CASES=r1n,gn python tools/ab.py r2h 3 build/libnewA10.so build/libnewA16.so build/libnewA10T.so > gpurun_out/r2h_ab.txt 2>&1
cat gpurun_out/r2h_noise.log gpurun_out/r2h_ab.txt
