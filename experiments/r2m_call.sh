#!/bin/bash
# round 2, GPU call M (8 GPUs): final code -- parity at W=8 (peer memory, forced copy engines, NCCL), c3 / c2 / c4 lines
set -u
OUT=gpurun_out/r2m
mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) "${@:2}"; }
timeout 400 bash -c "$(declare -f run); run 8 tests/dist_parity.py" > $OUT/parity_w8.log 2>&1; echo "parity W=8 rc=$?"; tail -2 $OUT/parity_w8.log
timeout 200 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 20 --warmup 5" > $OUT/bench_c3_w8.log 2>&1; echo "bench c3 W=8 rc=$?"
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_c3_w1.log 2>&1; echo "bench c3 W=1 rc=$?"
for c in c2 c4; do
  timeout 200 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --config $c --steps 20 --warmup 5" > $OUT/bench_${c}_w8.log 2>&1; echo "bench $c W=8 rc=$?"
done
for f in $OUT/bench_*.log; do echo $f; tail -1 $f | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["host_issue_ms_per_step"], d.get("parity",{}).get("ok"), d["op_ms_per_step"])' 2>&1 | tail -1; done
