#!/bin/bash
# A/B of the in-tree library (main) and scripts/dev/alt on the headline batch
mkdir -p gpurun_out
cp miniframe_b200/csrc/libminiframe_b200.so /tmp/main.so
for rep in 1 2 3; do
for lib in main alt; do
  if [ $lib = alt ]; then cp scripts/dev/alt/libminiframe_b200.so miniframe_b200/csrc/libminiframe_b200.so; else cp /tmp/main.so miniframe_b200/csrc/libminiframe_b200.so; fi
  timeout 300 python bench.py --steps 4 --warmup 2 --batch ${BATCH:-4096} --epochs ${EPOCHS:-500} --no-cpu-baseline > gpurun_out/bench_ab.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print("rep $rep $lib: build ms %.3f potrf ms %.2f (%.2f TF/s) value %.0f power %s" % (d["roofline_build"]["ms_per_launch"], d['roofline']['ms_per_launch'], d['roofline']['achieved'], d['value'], d['clocks']['power_w_max']))
PY
done
done
cp /tmp/main.so miniframe_b200/csrc/libminiframe_b200.so
