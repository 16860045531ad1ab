#!/bin/bash
# First GPU call of the next round: run the staged, default-off paths once each, cheapest first.
#   1 GPU :  gpurun --timeout 1500 -- 'bash experiments/validate_staged.sh 1'
#   N GPUs:  gpurun --gpus 2 --timeout 1500 -- 'bash experiments/validate_staged.sh 2'
# Everything lands in gpurun_out/staged_*.log; nothing here runs under a profiler.
set -u
N=${1:-1}
OUT=gpurun_out
mkdir -p $OUT
if [ "$N" = "1" ]; then
  MRCLIP_TEST_GRAPH=1 MRCLIP_TEST_FWDDS=1 timeout 600 python -m pytest tests/test_gpu_graph.py tests/test_gpu_fwdds.py -m gpu -q \
      > $OUT/staged_tests_n1.log 2>&1; echo "opt-in tests rc=$?" | tee -a $OUT/staged_tests_n1.log
  timeout 600 python bench.py --graph --steps 20 --warmup 3 --no-cpu-baseline > $OUT/staged_bench_graph_n1.log 2>&1
  tail -1 $OUT/staged_bench_graph_n1.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("eager ms", d["ms_per_step"], "graph", d.get("graph"))'
else
  NAMES=$(python - <<PY
import sys; sys.path.insert(0, "tests")
from conftest import golden_names, load_golden
print(" ".join(n for n in golden_names() if load_golden(n)["world"] == $N))
PY
)
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29655"
  MRCLIP_TEST_FWDDS=1 MRCLIP_TEST_PUSHBF16=1 timeout 900 $RUN tests/dist_worker.py $NAMES > $OUT/staged_dist_n$N.log 2>&1
  echo "dist_worker rc=$?" | tee -a $OUT/staged_dist_n$N.log
  for knobs in "" "MRCLIP_DS=fwd" "MRCLIP_PUSH_DTYPE=bf16" "MRCLIP_DS=fwd MRCLIP_PUSH_DTYPE=bf16"; do
    tag=$(echo "${knobs:-default}" | tr ' =' '__')
    env $knobs timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 3 > $OUT/staged_bench_n${N}_$tag.log 2>&1
    tail -1 $OUT/staged_bench_n${N}_$tag.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], d["ms_per_step"], d["op_ms_per_step"])' "$tag"
  done
  env timeout 600 $RUN bench.py --gpus $N --graph --steps 20 --warmup 3 > $OUT/staged_bench_n${N}_graph.log 2>&1
  tail -1 $OUT/staged_bench_n${N}_graph.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("graph", d["ms_per_step"], d.get("graph"))'
fi
