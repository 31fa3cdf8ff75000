#!/bin/bash
bash scripts/gpu_iter2.sh
echo "== 1 CTA/SM"
MF_SLOTS_PER_SM=1 timeout 300 python bench.py --steps 3 --warmup 2 --batch 1184 --no-cpu-baseline > gpurun_out/bench_sw.json 2>/dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/bench_sw.json'))
print("1 CTA/SM: potrf ms %.2f (%.2f TF/s)" % (d['roofline']['ms_per_launch'], d['roofline']['achieved']))
PY
