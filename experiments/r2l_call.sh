#!/bin/bash
# round 2, GPU call L (4 GPUs): three copy-engine lanes; smoke(); parity at W=4
set -u
OUT=gpurun_out/r2l
mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) "${@:2}"; }
timeout 200 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 300 bash -c "$(declare -f run); run 4 tests/dist_parity.py c3 c2 c4 c2raw mpos exchange" > $OUT/parity_w4.log 2>&1; echo "parity W=4 rc=$?"; tail -2 $OUT/parity_w4.log
for W in 4 2; do
  timeout 200 bash -c "$(declare -f run); run $W bench.py --gpus $W --steps 20 --warmup 5" > $OUT/bench_c3_w$W.log 2>&1; echo "bench c3 W=$W rc=$?"
done
MRCLIP_AG_OVERLAP=0 timeout 200 bash -c "$(declare -f run); run 4 bench.py --gpus 4 --steps 20 --warmup 5" > $OUT/bench_c3_w4_noov.log 2>&1; echo "bench c3 W=4 noov rc=$?"
timeout 200 bash -c "$(declare -f run); run 4 bench.py --gpus 4 --config c4 --steps 20 --warmup 5" > $OUT/bench_c4_w4.log 2>&1; echo "bench c4 W=4 rc=$?"
for f in $OUT/bench_*.log; do echo $f; tail -1 $f | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["host_issue_ms_per_step"], d.get("parity",{}).get("ok"), d["op_ms_per_step"])' 2>&1 | tail -1; done
