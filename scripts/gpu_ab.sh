#!/bin/bash
# A/B timings of the kernels (product, experiment and round-1-potf2 builds): correctness of the new kernels + A/B timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/ab_nvsmi.txt
echo "== A/B potrf + cov (product lib)"; timeout 600 python scripts/ab_kernels.py --cov > gpurun_out/ab_ab_product.log 2>&1; echo "rc=$?"; cat gpurun_out/ab_ab_product.log | tail -20
echo "== experiments (exp lib)"; timeout 600 python scripts/ab_kernels.py --lib libminiframe_b200_exp.so --exp 0,32,64,96,128,224,736,8,256 > gpurun_out/ab_ab_exp.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/ab_ab_exp.log
echo "== old potf2 (old lib)"; timeout 600 python scripts/ab_kernels.py --lib libminiframe_b200_old.so --exp 0,32,96 > gpurun_out/ab_ab_old.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/ab_ab_old.log
echo "== small n"; for N in 100 200; do timeout 300 python scripts/ab_kernels.py --epochs $N --batch 2664 2>&1 | tail -1; timeout 300 python scripts/ab_kernels.py --lib libminiframe_b200_old.so --epochs $N --batch 2664 2>&1 | tail -1; done > gpurun_out/ab_small.log 2>&1; cat gpurun_out/ab_small.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --durations=8 > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/ab_pytest.log | cut -c1-250
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err; echo "rc=$?"; tail -3 gpurun_out/ab_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/ab_bench.json'))
print("value %.0f e2e %.0f | potrf %.2f TF/s frac %.3f ms %.2f | build %.0f GB/s frac %.3f ms %.2f | peaks %s | clocks %s | cpu %s" % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline_build']['achieved'], d['roofline_build']['frac'], d['roofline_build']['ms_per_launch'], d['fp64_peaks_measured'], d['clocks'], d['cpu_baseline']))
PY
