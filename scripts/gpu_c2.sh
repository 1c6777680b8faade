#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/bench_c2.py --graph transport_Vermont --check 12 > gpurun_out/bench_c2_vermont.json 2> gpurun_out/bench_c2_vermont.err; echo "rc $?"; cat gpurun_out/bench_c2_vermont.json; tail -5 gpurun_out/bench_c2_vermont.err
