#!/bin/bash
# scripts/dev/ab_libs.sh OUT "exp codes" lib1 lib2 ...   (names without the libminiframe_b200_ prefix)
OUT=$1; EXP=$2; shift 2
mkdir -p gpurun_out
: > gpurun_out/$OUT
for v in "$@"; do
  lib=libminiframe_b200_$v.so; [ "$v" = product ] && lib=libminiframe_b200.so
  echo "== $v" >> gpurun_out/$OUT
  timeout ${ABTO:-90} python scripts/ab_kernels.py --lib $lib --exp $EXP --reps 3 2>&1 | grep -E "potrf exp|Error|error|finite" >> gpurun_out/$OUT
done
cat gpurun_out/$OUT
