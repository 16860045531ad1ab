#!/bin/bash
# round 2, GPU call B (2 GPUs): whole-step C entries + peer flags: GPU tests, production-size parity W=1/2, bench W=1/2
set -u
OUT=gpurun_out/r2b
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/gpu_tests.log 2>&1
echo "gpu tests rc=$?" | tee -a $OUT/gpu_tests.log
tail -4 $OUT/gpu_tests.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29655"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_n1.log 2>&1; echo "bench n1 rc=$?"; tail -1 $OUT/bench_n1.log | cut -c1-300
timeout 300 $RUN bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.log 2>&1; echo "bench n2 rc=$?"; tail -1 $OUT/bench_n2.log | cut -c1-300
MRCLIP_STEP=py timeout 300 $RUN bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2_py.log 2>&1; echo "bench n2 py rc=$?"; tail -1 $OUT/bench_n2_py.log | cut -c1-300
for f in $OUT/bench_n1.log $OUT/bench_n2.log $OUT/bench_n2_py.log; do tail -1 $f | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["host_issue_ms_per_step"], d["gpu_launches"], d["op_ms_per_step"])'; done
