#!/bin/bash
# Diagnostic: rebuild the library with parts of the covariance kernel switched off (PSI_LAB bits: 1 no DMMA,
# 2 no operand prefetch, 4 no staging stores) into tools/bin/, to see which part keeps the kernel off the
# store ceiling.  Run on the GPU box:  for v in 0 1 2 4 7; do FABLE_B200_LIB=tools/bin/libpsi_lab$v.so python tools/prof_psi.py ...; done
set -e
cd "$(dirname "$0")/../fable_b200/csrc"
mkdir -p ../../tools/bin/lab
for v in ${LABS:-0 1 2 3 4 7}; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-pthread -DPSI_LAB=$v -c psi.cu -o ../../tools/bin/lab/psi$v.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/bin/libpsi_lab$v.so build/api.o ../../tools/bin/lab/psi$v.o build/sampler.o build/rows.o build/reduce.o build/gemm.o build/eig.o build/svd.o build/post.o -lcudart -lpthread
done
