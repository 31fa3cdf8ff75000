#!/bin/bash
# A/B of two builds of the library on one box: the in-tree one and scripts/dev/alt/ (swapped in place)
mkdir -p gpurun_out
cp miniframe_b200/csrc/libminiframe_b200.so /tmp/main.so
for rep in 1 2; do
for lib in main alt; do
  if [ $lib = alt ]; then cp scripts/dev/alt/libminiframe_b200.so miniframe_b200/csrc/libminiframe_b200.so; else cp /tmp/main.so miniframe_b200/csrc/libminiframe_b200.so; fi
  for cfg in "$@"; do
    env $cfg timeout 300 python bench.py --steps 4 --warmup 2 --batch 2368 --no-cpu-baseline > gpurun_out/bench_ab.json 2>/dev/null
    python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print("rep $rep $lib [$cfg]: potrf ms %.2f (%.2f TF/s)" % (d['roofline']['ms_per_launch'], d['roofline']['achieved']))
PY
  done
done
done
cp /tmp/main.so miniframe_b200/csrc/libminiframe_b200.so
