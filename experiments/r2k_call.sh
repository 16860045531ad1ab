#!/bin/bash
# round 2, GPU call K (1 GPU): forward epilogue with independent max / sum chains
set -u
OUT=gpurun_out/r2k
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $OUT/gpu_tests.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_n1_$i.log 2>&1; echo "bench rc=$?"
tail -1 $OUT/bench_n1_$i.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["op_ms_per_step"])'
done
