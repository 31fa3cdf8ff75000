#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 4 7 8; do
  MF_SLOTS_PER_SM=1 MF_DBG=$dbg timeout 300 python bench.py --steps 3 --warmup 2 --batch 1184 --no-cpu-baseline > gpurun_out/bench_sw.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_sw.json'))
print("1 CTA/SM dbg $dbg: potrf ms %.2f" % (d['roofline']['ms_per_launch']))
PY
done
