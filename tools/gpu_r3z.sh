#!/bin/bash
# end-of-round check on one GPU: the whole GPU suite, the bench line (pipelined + serial e2e), the reference arm, and the
# half-height box that the 8-GPU e2e check used (profiles/r3e_bench_n8_nz256.json) for a like-for-like efficiency
o=gpurun_out; tag=r3z
BFLBM_STATS_OUT=$o timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $o/${tag}_pytest_gpu.log; cat $o/${tag}_pytest_gpu.log
timeout 300 python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; tail -c 400 $o/${tag}_bench_n1.json; echo
timeout 120 python bench.py --impl reference > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_ref.err
timeout 200 python bench.py --nz 256 --no-cpu --steps 10 > $o/${tag}_bench_n1_nz256.json 2>/dev/null; tail -c 300 $o/${tag}_bench_n1_nz256.json; echo
