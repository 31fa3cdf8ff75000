#!/bin/bash
# round 2, call C: dataflow cooperative factorisation - correctness, soak, timings
mkdir -p gpurun_out
echo "== pytest gpu (coop + configs)"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "cooperative or config or streams or small_tile or mixes or interleaved" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2c_pytest.log | cut -c1-250
echo "== soak"; timeout 900 python scripts/soak.py > gpurun_out/r2c_soak.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2c_soak.log
echo "== large / small batches"; timeout 900 python scripts/small_batch.py big > gpurun_out/r2c_small_batch.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2c_small_batch.log
echo "== mcmc demo"; timeout 600 python scripts/mcmc_demo.py > gpurun_out/r2c_mcmc.log 2>&1; cat gpurun_out/r2c_mcmc.log
echo "== n=1500 big batch unchanged?"; timeout 300 python scripts/ab_kernels.py 2>&1 | tail -1
