#!/bin/bash
# round 2, GPU call H (1 GPU): ncu evidence for the shipped step (launch list + full-set captures of the top kernels)
set -u
OUT=gpurun_out/r2h
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity"
$CMD > $OUT/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 4 -c 1 -o $OUT/fwd_tiles $CMD > $OUT/ncu_fwd.log 2>&1
echo "fwd capture rc=$?"
$CMD > $OUT/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 8 -c 2 -o $OUT/gemm2 $CMD > $OUT/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
$CMD > $OUT/plain4.log 2>&1 &&
ncu --set full --clock-control none -k regex:emat_transform -s 4 -c 1 -o $OUT/transform $CMD > $OUT/ncu_xf.log 2>&1
echo "transform capture rc=$?"
ls -la $OUT
