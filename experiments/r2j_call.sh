#!/bin/bash
# round 2, GPU call J (8 GPUs): final state -- production-size parity at W=8, scaling bench with the copy-engine all-gather
set -u
OUT=gpurun_out/r2j
mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) "${@:2}"; }
timeout 400 bash -c "$(declare -f run); run 8 tests/dist_parity.py" > $OUT/parity_w8.log 2>&1; echo "parity W=8 rc=$?"; tail -2 $OUT/parity_w8.log
for W in 8 4 2; do
  timeout 200 bash -c "$(declare -f run); run $W bench.py --gpus $W --steps 20 --warmup 5" > $OUT/bench_c3_w$W.log 2>&1; echo "bench c3 W=$W rc=$?"
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_c3_w1.log 2>&1; echo "bench c3 W=1 rc=$?"
MRCLIP_AG_OVERLAP=0 timeout 200 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 20 --warmup 5" > $OUT/bench_c3_w8_noov.log 2>&1; echo "bench c3 W=8 no-overlap rc=$?"
for c in c2 c4; do
  timeout 200 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --config $c --steps 20 --warmup 5" > $OUT/bench_${c}_w8.log 2>&1; echo "bench $c W=8 rc=$?"
done
timeout 300 bash -c "$(declare -f run); run 8 experiments/train_step.py --batch-per-gpu 1024 --steps 4 --warmup 2 --grad-checkpointing --out $OUT/train_step_n8_b1024.json" > $OUT/train_step_n8.log 2>&1; echo "train step W=8 rc=$?"
for f in $OUT/bench_*.log; do echo $f; tail -1 $f | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["host_issue_ms_per_step"], d["gpu_launches"], d.get("parity",{}).get("ok"), d["op_ms_per_step"])' 2>&1 | tail -1; done
tail -1 $OUT/train_step_n8.log | cut -c1-600
