#!/bin/bash
# round 2, call B: loop variants A/B, large-matrix regime, first ncu capture of the new kernels
mkdir -p gpurun_out
for lib in libminiframe_b200.so libminiframe_b200_e1.so libminiframe_b200_e2.so libminiframe_b200.so libminiframe_b200_e1.so libminiframe_b200_e2.so; do
  timeout 300 python scripts/ab_kernels.py --lib $lib --reps 4 2>&1 | tail -1 | sed "s/^/$lib: /"
done > gpurun_out/r2b_variants.log 2>&1; cat gpurun_out/r2b_variants.log
echo "== exp e2: 96 (loops only)"; timeout 300 python scripts/ab_kernels.py --lib libminiframe_b200_e2.so --exp 0,96,32 2>&1 | tail -3
echo "== cov"; timeout 300 python scripts/ab_kernels.py --cov 2>&1 | grep "mode 2" > gpurun_out/r2b_cov.log; cat gpurun_out/r2b_cov.log
echo "== n=6000 B=296"; timeout 600 python scripts/ab_kernels.py --epochs 2000 --batch 296 --reps 2 --cov 2>&1 | grep -v "mode 0" | tee gpurun_out/r2b_n6000.log
echo "== large / small batches"; timeout 900 python scripts/small_batch.py big > gpurun_out/r2b_small_batch.log 2>&1; tail -15 gpurun_out/r2b_small_batch.log
echo "== mcmc demo"; timeout 600 python scripts/mcmc_demo.py > gpurun_out/r2b_mcmc.log 2>&1; cat gpurun_out/r2b_mcmc.log
echo "== ncu full"; timeout 300 python scripts/prof_target.py 296 500 1 > gpurun_out/r2b_prof_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:mf_ -c 3 -o gpurun_out/prof_r2b -f python scripts/prof_target.py 296 500 1 > gpurun_out/r2b_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/r2b_ncu_full.log
