#!/bin/bash
# round 2, call H: full validation + profiles of the final kernels
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=5 > gpurun_out/val_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/val_pytest.log | cut -c1-250
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err; echo "rc=$?"; tail -3 gpurun_out/val_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/val_bench.json'))
print("value %.0f e2e %.0f | potrf %.2f TF/s frac %.3f ms %.2f | build %.0f GB/s frac %.3f ms %.2f | peaks %s | clocks %s | cpu %s" % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline_build']['achieved'], d['roofline_build']['frac'], d['roofline_build']['ms_per_launch'], d['fp64_peaks_measured'], d['clocks'], d['cpu_baseline']))
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/val_bench_ref.json 2> gpurun_out/val_bench_ref.err; echo "rc=$?"; cut -c1-600 gpurun_out/val_bench_ref.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/val_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/val_ncu_list.log 2>&1; echo "rc=$?"; grep -c "mf_" gpurun_out/val_launches.csv
echo "== ncu full"; timeout 1500 ncu --set full --clock-control none --import-source on -k regex:mf_ -c 3 -o gpurun_out/prof_val -f python scripts/prof_target.py 296 500 1 > gpurun_out/val_ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/val_ncu_full.log
echo "== fp64 probes"; timeout 300 python scripts/probe_fp64.py > gpurun_out/val_fp64_probes.txt 2>&1; tail -8 gpurun_out/val_fp64_probes.txt
