#!/bin/bash
# usage (2-GPU box): tools/gpu_r3b.sh -- pipelined e2e leg on slabs (peer stores and NCCL), driver tests with the writer thread
o=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu > $o/r3b_bench_n2_peer.json 2> $o/r3b_bench_n2_peer.err; tail -c 1800 $o/r3b_bench_n2_peer.json; tail -3 $o/r3b_bench_n2_peer.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu --halo nccl --nz 128 --e2e-steps 50 > $o/r3b_bench_n2_nccl_small.json 2> $o/r3b_bench_n2_nccl.err; tail -c 1200 $o/r3b_bench_n2_nccl_small.json; tail -3 $o/r3b_bench_n2_nccl.err
timeout 600 python -m pytest tests/test_host_driver.py tests/test_gpu_multiprocess.py -m gpu -q 2>&1 | tail -8 > $o/r3b_pytest_driver_mp.log; cat $o/r3b_pytest_driver_mp.log
