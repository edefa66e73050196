#!/bin/bash
# 8 GPUs, 512x512x256 cells per GPU (half the bench height: halves every host copy AND the steps, same upload : steps ratio):
# serial and pipelined e2e on the slab path, weak scaling
o=gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --no-cpu --nz 256 --steps 10 > $o/r3e_bench_n8_nz256.json 2> $o/r3e_bench_n8.err; tail -c 1500 $o/r3e_bench_n8_nz256.json; tail -3 $o/r3e_bench_n8.err
