#!/bin/bash
# 2-GPU validation of the native multi-GPU paths
export BFLBM_MP_LOG=$PWD/gpurun_out/r2e_mp_slab.log
python -m pytest tests/test_gpu_multiprocess.py tests/test_reference_numbers.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2e_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2e_bench_n2_peer.json 2> gpurun_out/r2e_bench_n2_peer.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e --halo nccl > gpurun_out/r2e_bench_n2_nccl.json 2> gpurun_out/r2e_bench_n2_nccl.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu --no-e2e --scaling strong --nz 128 > gpurun_out/r2e_bench_n2_strong128_peer.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu --no-e2e --scaling strong --nz 128 --halo nccl > gpurun_out/r2e_bench_n2_strong128_nccl.json 2>/dev/null
cat gpurun_out/r2e_pytest.log gpurun_out/r2e_mp_slab.log; tail -c 600 gpurun_out/r2e_bench_n2_peer.err
for f in gpurun_out/r2e_bench_n2_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], '%.0f MLUPS %.3f ms' % (d['value'], d['ms_per_step']), d['config'].get('slab_parity'), d['config'].get('halo','')[:20], 'e2e', d['e2e'] and round(d['e2e']['value']))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
