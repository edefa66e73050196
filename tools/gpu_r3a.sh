#!/bin/bash
# usage (GPU box): tools/gpu_r3a.sh -- asynchronous host transfers: tests, then the bench line with the pipelined e2e leg
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_async_transfers.py -q -x 2>&1 | tail -25 > $o/r3a_async_tests.log; cat $o/r3a_async_tests.log
timeout 600 python bench.py > $o/r3a_bench_n1.json 2> $o/r3a_bench_n1.err; tail -c 2500 $o/r3a_bench_n1.json; tail -5 $o/r3a_bench_n1.err
