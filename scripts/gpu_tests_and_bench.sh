#!/bin/bash
# tests + quick bench
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench 592"; timeout 600 python bench.py --steps 3 --warmup 2 --batch 592 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_small.json'))
print("value %.0f e2e %.0f | potrf %.2f TF/s frac %.3f ms %.2f | build %.0f GB/s frac %.3f ms %.2f | peaks %s | clocks %s" % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline_build']['achieved'], d['roofline_build']['frac'], d['roofline_build']['ms_per_launch'], d['fp64_peaks_measured'], d['clocks']))
PY
tail -3 gpurun_out/bench_small.err
echo "== bench full"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print("value %.0f e2e %.0f | potrf %.2f TF/s frac %.3f ms %.2f | build %.0f GB/s frac %.3f ms %.2f | clocks %s" % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline_build']['achieved'], d['roofline_build']['frac'], d['roofline_build']['ms_per_launch'], d['clocks']))
PY
tail -3 gpurun_out/bench_full.err
