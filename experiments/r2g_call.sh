#!/bin/bash
# round 2, GPU call G (1 GPU): stream-K gradient GEMMs
set -u
OUT=gpurun_out/r2g
mkdir -p $OUT
timeout 600 python tests/dist_parity.py c3 c2 c2raw c4 mpos ragged > $OUT/parity_w1.log 2>&1; echo "parity rc=$?"; grep -c "\[ok\]" $OUT/parity_w1.log; grep -E "FAIL|MISMATCH|mrclip:" $OUT/parity_w1.log | head -5
for k in 1 0; do
  MRCLIP_STREAMK=$k timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_sk$k.log 2>&1; echo "streamk=$k rc=$?"
  tail -1 $OUT/bench_sk$k.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["op_ms_per_step"])'
done
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $OUT/gpu_tests.log
