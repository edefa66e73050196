#!/bin/bash
# end-of-round check on one GPU: the whole GPU suite (16-run statistical ensembles against the regenerated goldens) and the
# bench lines with the kernel duration taken inside the timed region
o=gpurun_out; tag=${1:-r2q}
BFLBM_STATS_OUT=$o python -m pytest tests -m gpu -q 2>&1 | tail -6 > $o/${tag}_pytest_gpu.log; cat $o/${tag}_pytest_gpu.log
python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; tail -c 600 $o/${tag}_bench_n1.json; echo
python bench.py --impl reference > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_ref.err
python bench.py --kbt 0 --no-e2e --no-cpu --steps 20 > $o/${tag}_bench_n1_det.json 2>/dev/null
BFLBM_RATE1=0 python bench.py --no-e2e --no-cpu --steps 20 > $o/${tag}_bench_n1_general.json 2>/dev/null
BFLBM_RATE1=0 python bench.py --kbt 0 --no-e2e --no-cpu --steps 20 > $o/${tag}_bench_n1_general_det.json 2>/dev/null
python bench.py --algo auto --nx 256 --ny 256 --nz 256 --steps 200 --warmup 20 --no-e2e --no-cpu > $o/${tag}_bench_256cubed_noise.json 2>/dev/null
ls -la $o | grep ${tag}
