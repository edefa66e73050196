#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_host_driver.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2d_parity.log
tools/smallbox2.sh r2d_smallbox > /dev/null 2>&1
AB_CASES=gn,gd,r1n python tools/ab.py r2d 2 build/libold.so build/libnewD.so build/libnewGE.so > gpurun_out/r2d_ab.txt 2>&1
cat gpurun_out/r2d_parity.log gpurun_out/r2d_smallbox.txt gpurun_out/r2d_ab.txt
