#!/bin/bash
# round 2, GPU call C (1 GPU): banded backward (rescale of band b+1 on a side stream under the dI GEMM of band b)
set -u
OUT=gpurun_out/r2c
mkdir -p $OUT
timeout 600 python tests/dist_parity.py c3 c2 c2raw c4 ragged > $OUT/parity_w1.log 2>&1; echo "parity rc=$?"; tail -3 $OUT/parity_w1.log
for b in 1 2 4 8 16; do
  MRCLIP_BANDS=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_bands$b.log 2>&1; echo "bands=$b rc=$?"
  tail -1 $OUT/bench_bands$b.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["op_ms_per_step"], d["clocks"])'
done
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $OUT/gpu_tests.log
