#!/bin/bash
# round 2, GPU call D (8 GPUs): multi-GPU parity at production sizes (W=8, W=4), goldens, scaling bench, configs c2/c4, graphs
set -u
OUT=gpurun_out/r2d
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) "${@:2}"; }
for W in 8 4; do
  timeout 400 bash -c "$(declare -f run); run $W tests/dist_parity.py" > $OUT/parity_w$W.log 2>&1; echo "parity W=$W rc=$?"; tail -2 $OUT/parity_w$W.log
done
NAMES=$(python - <<PY
import sys; sys.path.insert(0, "tests")
from conftest import golden_names, load_golden
print(" ".join(n for n in golden_names() if load_golden(n)["world"] == 8))
PY
)
timeout 400 bash -c "$(declare -f run); run 8 tests/dist_worker.py $NAMES" > $OUT/goldens_w8.log 2>&1; echo "goldens W=8 rc=$?"; tail -3 $OUT/goldens_w8.log
for W in 8 4 2; do
  timeout 200 bash -c "$(declare -f run); run $W bench.py --gpus $W --steps 20 --warmup 5" > $OUT/bench_c3_w$W.log 2>&1; echo "bench c3 W=$W rc=$?"
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_c3_w1.log 2>&1; echo "bench c3 W=1 rc=$?"
for c in c2 c4; do
  timeout 200 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --config $c --steps 20 --warmup 5" > $OUT/bench_${c}_w8.log 2>&1; echo "bench $c W=8 rc=$?"
done
timeout 150 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --config c2 --graph --steps 20 --warmup 5" > $OUT/bench_c2_w8_graph.log 2>&1; echo "bench c2 graph W=8 rc=$?"
timeout 150 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --graph --steps 20 --warmup 5" > $OUT/bench_c3_w8_graph.log 2>&1; echo "bench c3 graph W=8 rc=$?"
for f in $OUT/bench_*.log; do echo $f; tail -1 $f | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["host_issue_ms_per_step"], d["gpu_launches"], d.get("graph"), d.get("parity",{}).get("worst_over_ranks"), d["op_ms_per_step"])' 2>&1 | tail -1; done
