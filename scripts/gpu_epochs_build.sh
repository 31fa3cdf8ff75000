#!/bin/bash
# covariance-build bandwidth as a function of N (alignment of the block offsets)
mkdir -p gpurun_out
for N in 496 500 512 528; do
  timeout 300 python bench.py --steps 4 --warmup 2 --batch 1184 --epochs $N --no-cpu-baseline > gpurun_out/bench_ab.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print("N $N: build ms %.3f (%.0f GB/s)  potrf %.2f TF/s" % (d['roofline_build']['ms_per_launch'], d['roofline_build']['achieved'], d['roofline']['achieved']))
PY
done
