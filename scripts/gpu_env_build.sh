#!/bin/bash
# covariance-build time for environment-selected launch shapes, batch 2368
mkdir -p gpurun_out
for rep in 1 2; do
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --steps 4 --warmup 2 --batch 2368 --no-cpu-baseline > gpurun_out/bench_ab.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print("rep $rep [$cfg]: build ms %.3f (%.0f GB/s)  value %.0f" % (d['roofline_build']['ms_per_launch'], d['roofline_build']['achieved'], d['value']))
PY
done
done
