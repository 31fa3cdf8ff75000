#!/bin/bash
# quick iteration: diagnostics + GPU tests (+ optional ncu)
mkdir -p gpurun_out
echo "== diag"; timeout 600 python scripts/diag_cov.py > gpurun_out/diag_cov.log 2>&1; echo "diag rc=$?"; grep -E "asym=[1-9]|nviol=[1-9]|^N=|rel=[1-9]\.[0-9]+e-(0|1[01])" gpurun_out/diag_cov.log | cut -c1-300
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
if [ "$1" == "ncu" ]; then
  echo "== ncu launch list"
  timeout 300 python scripts/prof_target.py 296 500 2 > gpurun_out/prof_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/prof_target.py 296 500 2 > gpurun_out/ncu_list.log 2>&1
  echo "ncu list rc=$?"; grep -E "mf_|ampere|cutlass" gpurun_out/launches.csv | cut -d, -f5,12- | head -20
  echo "== ncu full"
  timeout 300 python scripts/prof_target.py 296 500 1 > gpurun_out/prof_plain2.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:mf_ -c 2 -o gpurun_out/prof_r1 -f python scripts/prof_target.py 296 500 1 > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -5 gpurun_out/ncu_full.log
fi
