#!/bin/bash
# general-rate kernel on the rate-1 schedule with a second fetch of the old populations (BFLBM_GENERAL_REREAD) against the shipped one
o=gpurun_out
AB_CASES=gn,gd python tools/ab.py r3c 2 build/libBASE.so build/libRR.so build/libRR4.so > $o/r3c_ab.txt 2>&1; cat $o/r3c_ab.txt
for v in RR RR4; do
  BFLBM_LIB=$PWD/build/lib$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 > $o/r3c_parity_$v.log; cat $o/r3c_parity_$v.log
done
