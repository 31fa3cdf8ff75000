#!/bin/bash
mkdir -p gpurun_out
echo "== grad + sampler tests"; timeout 900 python -m pytest tests/test_grad.py tests/test_sampler.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/r2f_grad.log 2>&1; echo "rc=$?"; grep -E "worst|passed|failed|Error|error" gpurun_out/r2f_grad.log | head -20; tail -12 gpurun_out/r2f_grad.log | cut -c1-220
echo "== mcmc demo"; timeout 600 python scripts/mcmc_demo.py > gpurun_out/r2f_mcmc.log 2>&1; cat gpurun_out/r2f_mcmc.log | cut -c1-300
