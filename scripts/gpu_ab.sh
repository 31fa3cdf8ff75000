#!/bin/bash
# A/B of kernel variants on the same box: MF_VARIANT=0 vs 1, alternating
mkdir -p gpurun_out
echo "== tests with variant 1"; MF_VARIANT=1 timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2 3; do
  for v in 0 1; do
    MF_VARIANT=$v timeout 300 python bench.py --steps 4 --warmup 2 --batch 2368 --no-cpu-baseline > gpurun_out/bench_ab.json 2>/dev/null
    python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print("rep $rep variant $v: potrf ms %.2f (%.2f TF/s) value %.0f" % (d['roofline']['ms_per_launch'], d['roofline']['achieved'], d['value']))
PY
  done
done
