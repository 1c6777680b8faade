#!/bin/bash
mkdir -p gpurun_out
KR_SCREEN_EXACT=1 python scripts/bench_screen.py > gpurun_out/r02v_bench_screen_c5.json 2> gpurun_out/r02v_bench_screen_c5.err; echo "rc $?"; cat gpurun_out/r02v_bench_screen_c5.json; tail -3 gpurun_out/r02v_bench_screen_c5.err
