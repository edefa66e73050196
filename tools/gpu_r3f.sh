#!/bin/bash
# 2-GPU box: the bench with the untimed staging probe (N = 2 full size, N = 1 small), and the fallback to the serial interval
o=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --no-cpu > $o/r3f_bench_n2.json 2> $o/r3f_bench_n2.err; tail -c 900 $o/r3f_bench_n2.json; tail -3 $o/r3f_bench_n2.err
timeout 200 python bench.py --nz 128 --no-cpu --steps 10 > $o/r3f_bench_n1_nz128.json 2> $o/r3f_n1.err; tail -c 700 $o/r3f_bench_n1_nz128.json; tail -3 $o/r3f_n1.err
BFLBM_BENCH_NO_STAGING=1 timeout 200 python bench.py --nz 128 --no-cpu --steps 10 > $o/r3f_bench_n1_fallback.json 2> $o/r3f_n1fb.err; tail -c 700 $o/r3f_bench_n1_fallback.json; tail -3 $o/r3f_n1fb.err
