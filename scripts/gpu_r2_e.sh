#!/bin/bash
mkdir -p gpurun_out
echo "== grad tests"; timeout 600 python -m pytest tests/test_grad.py -m gpu -q --tb=short -p no:cacheprovider -x -s > gpurun_out/r2e_grad.log 2>&1; echo "rc=$?"; grep -E "worst|passed|failed|Error|error" gpurun_out/r2e_grad.log | head -20; tail -15 gpurun_out/r2e_grad.log | cut -c1-220
echo "== mapping"; timeout 900 python scripts/small_batch.py big > gpurun_out/r2e_small_batch.log 2>&1; echo "rc=$?"; cat gpurun_out/r2e_small_batch.log
