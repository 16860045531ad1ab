#!/bin/bash
# round 2, GPU call E (2 GPUs): overlapped text push (side stream, per-destination flags) + retrieval metrics + N2 train step
set -u
OUT=gpurun_out/r2e
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 $OUT/gpu_tests.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29655"
timeout 200 $RUN bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.log 2>&1; echo "bench n2 rc=$?"
MRCLIP_AG_OVERLAP=0 timeout 200 $RUN bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2_noov.log 2>&1; echo "bench n2 no-overlap rc=$?"
for f in $OUT/bench_n2.log $OUT/bench_n2_noov.log; do tail -1 $f | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["host_issue_ms_per_step"], d["gpu_launches"], d["op_ms_per_step"])'; done
timeout 600 python experiments/train_step.py --batch-per-gpu 512 --steps 4 --warmup 2 --out $OUT/train_step_n1_b512.json > $OUT/train_step_n1.log 2>&1; echo "train step rc=$?"; tail -1 $OUT/train_step_n1.log | cut -c1-900
timeout 600 $RUN experiments/train_step.py --batch-per-gpu 1024 --steps 4 --warmup 2 --grad-checkpointing --out $OUT/train_step_n2_b1024.json > $OUT/train_step_n2.log 2>&1; echo "train step W=2 rc=$?"; tail -1 $OUT/train_step_n2.log | cut -c1-900
