#!/bin/bash
# round 2, GPU call I (1 GPU): MultiPositiveClipLoss with the class-mean kernels
set -u
OUT=gpurun_out/r2i
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $OUT/gpu_tests.log
timeout 300 python bench.py --workload mpos --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_mpos.log 2>&1; echo "mpos rc=$?"
tail -1 $OUT/bench_mpos.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d.get("parity",{}).get("worst_over_ranks"), d["op_ms_per_step"])'
timeout 300 python bench.py --workload siglip --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_siglip.log 2>&1; echo "siglip rc=$?"
tail -1 $OUT/bench_siglip.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d.get("parity",{}).get("worst_over_ranks"), d["op_ms_per_step"])'
