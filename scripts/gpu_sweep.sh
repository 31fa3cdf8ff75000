#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 32 64 96 128 224 8; do
  MF_DBG=$dbg timeout 300 python bench.py --steps 3 --warmup 2 --batch 2368 --no-cpu-baseline > gpurun_out/bench_sw.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_sw.json'))
print("dbg $dbg: potrf ms %.2f" % (d['roofline']['ms_per_launch']))
PY
done
