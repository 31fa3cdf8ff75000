#!/bin/bash
# One gpurun call: smoke, GPU parity tests, a short and a full bench run.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== bench small"; timeout 600 python bench.py --steps 2 --warmup 1 --batch 592 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench small rc=$?"; cat gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
echo "== bench full"; timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
