#!/bin/bash
# general-rate kernel: species f by cp.async through shared memory (one DRAM latency per plane) against the sequential loads
o=gpurun_out
for v in FA0 FA3; do
  BFLBM_LIB=$PWD/build/lib$v.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 > $o/r2o_parity_$v.log; cat $o/r2o_parity_$v.log
done
AB_CASES=gn,gd python tools/ab.py r2o 2 build/libBASE.so build/libFA0.so build/libFA3.so > $o/r2o_ab.txt 2>&1; cat $o/r2o_ab.txt
BFLBM_CARVEOUT=100 AB_CASES=gn python tools/ab.py r2o_c100 1 build/libFA0.so build/libFA3.so > $o/r2o_ab_c100.txt 2>&1; cat $o/r2o_ab_c100.txt
