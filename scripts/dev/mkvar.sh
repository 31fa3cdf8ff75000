#!/bin/bash
# A/B build of the factorisation only: scripts/dev/mkvar.sh NAME "-DMF_..."  ->  libminiframe_b200_NAME.so
# (mf_potrf.cu compiled with -DMF_EXPERIMENTS and the given flags, linked with the other objects of `make exp`)
set -e
cd "$(dirname "$0")/../../miniframe_b200/csrc"
V=$1; D=$2
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v -DMF_EXPERIMENTS $D \
    -c mf_potrf.cu -o mf_potrf.$V.o 2> mf_potrf.$V.ptxas.log || { cat mf_potrf.$V.ptxas.log; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libminiframe_b200_$V.so mf_potrf.$V.o \
    mf_api.exp.o mf_cov.exp.o mf_solve.exp.o mf_grad.exp.o mf_means.exp.o mf_probe.exp.o
grep -A2 "ILi3ELi2ELi0ELi0ELi2" mf_potrf.$V.ptxas.log | grep -E "spill|Used" | tr '\n' ' '; echo
