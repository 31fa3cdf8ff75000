#!/bin/bash
mkdir -p gpurun_out
for st in 0 100000 1000000 -100000 -1000000 -4000000; do
  MF_STAGGER=$st timeout 300 python bench.py --steps 3 --warmup 2 --batch 1184 --no-cpu-baseline > gpurun_out/bench_sw.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_sw.json'))
print("stagger $st: potrf ms %.2f  (%.2f TF/s)" % (d['roofline']['ms_per_launch'], d['roofline']['achieved']))
PY
done
