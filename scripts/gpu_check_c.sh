#!/bin/bash
# tests + bench + ncu (launch list and one full capture of the SpMM) for the current build
TAG=${1:-c}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
tail -12 gpurun_out/pytest_$TAG.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_$TAG.log
tail -3 gpurun_out/bench_$TAG.log
export KR_BENCH_K=64 KR_BENCH_M=4
python bench.py --steps 1 --warmup 3 > gpurun_out/bench_small_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_launch_$TAG.log 2>&1
python bench.py --steps 1 --warmup 3 > gpurun_out/bench_small2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -s 12 -c 2 -o gpurun_out/prof_spmm_$TAG \
    python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_$TAG.log 2>&1
ls gpurun_out | head -40
