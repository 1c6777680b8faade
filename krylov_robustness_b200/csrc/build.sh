#!/bin/bash
# Build libkrylov_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC,-Wall,-Wno-unused-function,-fopenmp,-pthread -shared \
    ${KR_PTXAS_V:+-Xptxas -v} ${KR_EXTRA_FLAGS} \
    -o ${KR_OUT:-../libkrylov_b200.so} api.cu \
    -lgomp -Xlinker -rpath=/usr/local/cuda/lib64
