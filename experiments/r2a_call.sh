#!/bin/bash
# round 2, GPU call A (2 GPUs): default + staged GPU tests, production-size parity at W=1/2, staged benches
set -u
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
MRCLIP_TEST_GRAPH=1 MRCLIP_TEST_FWDDS=1 timeout 1200 python -m pytest tests -m gpu -q > $OUT/gpu_tests.log 2>&1
echo "gpu tests rc=$?" | tee -a $OUT/gpu_tests.log
tail -5 $OUT/gpu_tests.log
# (then: the opt-in tests and bench variants of the round-1 staged paths, experiments/validate_staged.sh of that commit)
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.log 2>&1; tail -1 $OUT/bench_n1.log | cut -c1-600
