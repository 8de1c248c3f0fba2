#!/bin/bash
# time the tensor-core kernels of the product build and of every experiment build under build/variants
mkdir -p gpurun_out
OUT=gpurun_out/${1:-r02_variants}.txt
: > $OUT
shopt -s nullglob
for lib in "" build/variants/*.so; do
  echo "== ${lib:-product}" >> $OUT
  SPN_LIB_PATH=${lib:+$PWD/$lib} timeout 300 python tests/_tc_time.py ${2:-all} > gpurun_out/_t.log 2>&1 || tail -5 gpurun_out/_t.log >> $OUT
  grep -E "^tc |^aborts|^prepare|^ +[0-9.]+ ms" gpurun_out/_t.log >> $OUT
done
cat $OUT
