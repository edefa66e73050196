#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2c_parity.log
tools/smallbox.sh r2c_smallbox > /dev/null 2>&1
AB_CASES=gn,gd python tools/ab.py r2c 2 build/libold.so build/libnewOLDW.so build/libnewOLDW2.so build/libnewFMGM.so build/libnewGMi.so > gpurun_out/r2c_ab.txt 2>&1
cat gpurun_out/r2c_parity.log gpurun_out/r2c_smallbox.txt gpurun_out/r2c_ab.txt
