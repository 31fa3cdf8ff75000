#!/bin/bash
# Round profile: launch list of the bench command + full ncu capture of both kernels (one wave).
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_for_ncu.json 2> gpurun_out/bench_for_ncu.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; grep -c "mf_" gpurun_out/launches_bench.csv
timeout 300 python scripts/prof_target.py 296 500 1 > gpurun_out/prof_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:mf_ -c 2 -o gpurun_out/prof_r1_final -f python scripts/prof_target.py 296 500 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
