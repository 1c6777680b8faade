#!/bin/bash
# one full ncu capture of the SpMM on the small bench configuration (same kernel, 8 panels)
TAG=${1:-p}
mkdir -p gpurun_out
export KR_BENCH_K=64 KR_BENCH_M=4
python bench.py --steps 1 --warmup 3 > gpurun_out/bench_small_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -s 12 -c 1 -o gpurun_out/prof_spmm_$TAG \
    python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/prof_spmm_$TAG.ncu-rep
