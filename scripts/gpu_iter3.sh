#!/bin/bash
bash scripts/gpu_iter2.sh
echo "== ncu cov"
timeout 300 python scripts/prof_target.py 592 500 1 > gpurun_out/prof_plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mf_cov -c 1 -o gpurun_out/prof_r1_cov -f python scripts/prof_target.py 592 500 1 > gpurun_out/ncu_cov.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_cov.log
