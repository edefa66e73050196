#!/bin/bash
# round 2, call b: general-rate kernel variants (parity of two of them, then interleaved A/B), carve-out sensitivity of the rate-1 kernel
for L in libnewGM libnewFMGMS; do
  BFLBM_LIB=$PWD/build/$L.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "injected or golden or rate1 or ragged" 2>&1 | tail -2 | sed "s/^/$L: /"
done > gpurun_out/r2b_parity.log
AB_CASES=gn,gd python tools/ab.py r2b 2 build/libold.so build/libnewG19.so build/libnewGM.so build/libnewGMS.so build/libnewS.so build/libnewFMGM.so build/libnewFMGMS.so > gpurun_out/r2b_ab.txt 2>&1
for c in 100 86 57; do
  BFLBM_CARVEOUT=$c AB_CASES=r1n,r1d python tools/ab.py r2b_carve$c 2 build/libnewG19.so 2>&1 | sed "s/^/carveout $c: /"
done > gpurun_out/r2b_carve.txt
cat gpurun_out/r2b_parity.log gpurun_out/r2b_ab.txt gpurun_out/r2b_carve.txt
