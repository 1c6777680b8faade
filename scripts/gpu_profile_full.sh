#!/bin/bash
# ncu capture of ONE full-size SpMM launch (k = 512) + launch list of a short full-size bench
TAG=${1:-full}
mkdir -p gpurun_out
export KR_BENCH_M=2 KR_BENCH_EDGES=0
python bench.py --steps 1 --warmup 3 > gpurun_out/bench_m2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -s 6 -c 1 -o gpurun_out/prof_spmm_$TAG \
    python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_$TAG.log 2>&1
python bench.py --steps 1 --warmup 3 > gpurun_out/bench_m2b_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_launch_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
