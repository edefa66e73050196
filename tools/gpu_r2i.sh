#!/bin/bash
# 8-GPU validation: real processes over NCCL and over peer-mapped mailboxes, bflbm_multi, the C++ driver, SF on slabs; bench lines
export BFLBM_MP_LOG=$PWD/gpurun_out/r2i_mp_slab.log
rm -f $BFLBM_MP_LOG $BFLBM_MP_LOG.full
python -m pytest tests/test_gpu_multiprocess.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r2i_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29811 bench.py --gpus 8 --no-cpu > gpurun_out/r2i_bench_n8_weak_peer.json 2> gpurun_out/r2i_bench_n8.err
$TR --nproc-per-node 8 --master-port 29812 bench.py --gpus 8 --no-cpu --no-e2e --scaling strong > gpurun_out/r2i_bench_n8_strong_peer.json 2>/dev/null
$TR --nproc-per-node 8 --master-port 29813 bench.py --gpus 8 --no-cpu --no-e2e --scaling strong --halo nccl > gpurun_out/r2i_bench_n8_strong_nccl.json 2>/dev/null
$TR --nproc-per-node 4 --master-port 29814 bench.py --gpus 4 --no-cpu > gpurun_out/r2i_bench_n4_weak_peer.json 2>/dev/null
cat gpurun_out/r2i_pytest.log $BFLBM_MP_LOG; tail -c 400 gpurun_out/r2i_bench_n8.err
for f in gpurun_out/r2i_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], '%.0f MLUPS %.3f ms' % (d['value'], d['ms_per_step']), d['config'].get('slab_parity'), (d['config'].get('halo') or '')[:12], 'e2e', d['e2e'] and (round(d['e2e']['value']), round(d['e2e']['seconds'],2)), d['config'].get('numa'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
