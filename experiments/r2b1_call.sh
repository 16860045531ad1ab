#!/bin/bash
# round 2, GPU call B1 (1 GPU): whole-step C entries: GPU tests, production-size parity W=1, bench
set -u
OUT=gpurun_out/r2b
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/gpu_tests_1gpu.log 2>&1
echo "gpu tests rc=$?" | tee -a $OUT/gpu_tests_1gpu.log
tail -4 $OUT/gpu_tests_1gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.log 2>&1; echo "bench n1 rc=$?"; tail -1 $OUT/bench_n1.log | cut -c1-300
MRCLIP_STEP=py timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_n1_py.log 2>&1; echo "bench n1 py rc=$?"
for c in c1 c2 c4; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 > $OUT/bench_n1_$c.log 2>&1; echo "bench $c rc=$?"; done
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 $OUT/bench_ref.log | cut -c1-400
for f in $OUT/bench_n1.log $OUT/bench_n1_py.log $OUT/bench_n1_c1.log $OUT/bench_n1_c2.log $OUT/bench_n1_c4.log; do tail -1 $f | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["host_issue_ms_per_step"], d["gpu_launches"], d.get("parity",{}).get("worst_over_ranks"), d["op_ms_per_step"], d.get("cpu_baseline"))'; done
