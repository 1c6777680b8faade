#!/bin/bash
# first GPU pass: parity tests of the SpMM/SLQ slice, smoke, bench, ncu launch list + one full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests/test_gpu_spmm.py -m gpu -x -q > gpurun_out/pytest_a.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_a.log
tail -5 gpurun_out/pytest_a.log
python __graft_entry__.py smoke > gpurun_out/smoke_a.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_a.log
tail -3 gpurun_out/smoke_a.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_a.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_a.log
tail -3 gpurun_out/bench_a.log
export KR_BENCH_K=64 KR_BENCH_M=4
python bench.py --steps 1 --warmup 3 > gpurun_out/bench_small_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_a.csv \
    python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_launch_a.log 2>&1
python bench.py --steps 1 --warmup 3 > gpurun_out/bench_small_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -s 12 -c 2 -o gpurun_out/prof_spmm_a \
    python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_a.log 2>&1
ls -la gpurun_out
