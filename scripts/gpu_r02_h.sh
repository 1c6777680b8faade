#!/bin/bash
# round 2, call H: the re-written bench.py end to end on 1 GPU (C3 headline + C5 + C4), reference arm
mkdir -p gpurun_out
KR_BENCH_C5_CAND=8192 KR_BENCH_C4_SCALE=20 python bench.py --steps 2 --warmup 3 > gpurun_out/r02h_bench_small.json 2> gpurun_out/r02h_bench_small.err; echo "small exit $?"; tail -c 600 gpurun_out/r02h_bench_small.err
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "full exit $?"; tail -5 gpurun_out/r02h_bench.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02h_bench_reference.json 2> gpurun_out/r02h_bench_reference.err; echo "ref exit $?"; tail -4 gpurun_out/r02h_bench_reference.err
python -m pytest tests/test_gpu_krylov.py -m gpu -q --timeout=900 -k "scaled_grid" 2>&1 | tail -3
