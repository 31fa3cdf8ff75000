#!/bin/bash
# sweep of environment-selected kernel knobs on one box: each line of $1.. is an "ENV=val ENV=val" set
mkdir -p gpurun_out
for rep in 1 2; do
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --steps 4 --warmup 2 --batch 2368 --no-cpu-baseline > gpurun_out/bench_ab.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print("rep $rep [$cfg]: potrf ms %.2f (%.2f TF/s) value %.0f" % (d['roofline']['ms_per_launch'], d['roofline']['achieved'], d['value']))
PY
done
done
