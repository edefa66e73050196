#!/bin/bash
# end-of-round check on two GPUs: slab parity over NCCL and peer stores, bflbm_multi, driver; bench line with e2e phases
o=gpurun_out; tag=${1:-r2q}
python -m pytest tests/test_gpu_multiprocess.py -m gpu -q 2>&1 | tail -5 > $o/${tag}_pytest_2gpu.log; cat $o/${tag}_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > $o/${tag}_bench_n2.json 2> $o/${tag}_bench_n2.err; tail -c 1500 $o/${tag}_bench_n2.json
