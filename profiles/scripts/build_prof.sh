#!/bin/bash
set -e
cd /root/repo/recombiner_b200/csrc
nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -DRCB_MLP_PROFILE -c mlp_tc.cu -o /root/repo/scratch/mlp_tc_prof.o
objs=$(ls build/*.o | grep -v "build/mlp_tc.o")
nvcc -shared -o /root/repo/scratch/libprof.so -gencode arch=compute_100a,code=sm_100a $objs /root/repo/scratch/mlp_tc_prof.o
echo built
