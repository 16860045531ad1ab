#!/bin/bash
# round 2, GPU call N (1 GPU): the committed tree as the driver will run it: GPU tests, smoke, default bench
set -u
OUT=gpurun_out/r2n
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q -x > $OUT/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 $OUT/gpu_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 200 python bench.py > $OUT/bench.log 2>&1; echo "bench rc=$?"; tail -1 $OUT/bench.log | cut -c1-1500
