#!/bin/bash
# round 2, GPU call F (2 GPUs): text all-gather on the copy engines, overlapped with the forward
set -u
OUT=gpurun_out/r2f
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29655"
timeout 200 $RUN bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.log 2>&1; echo "bench n2 rc=$?"
MRCLIP_AG_OVERLAP=0 timeout 200 $RUN bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2_noov.log 2>&1; echo "bench n2 no-overlap rc=$?"
for f in $OUT/bench_n2.log $OUT/bench_n2_noov.log; do tail -1 $f | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["host_issue_ms_per_step"], d["gpu_launches"], d["op_ms_per_step"])'; done
grep -m3 "mrclip:" $OUT/bench_n2.log
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 $OUT/gpu_tests.log
