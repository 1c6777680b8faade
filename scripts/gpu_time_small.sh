#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/time_small.py grid_England transport_Rome oregon_A8 > gpurun_out/time_small.jsonl 2> gpurun_out/time_small.err; echo "rc $?"; cat gpurun_out/time_small.jsonl; tail -5 gpurun_out/time_small.err
