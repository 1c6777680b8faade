#!/bin/bash
mkdir -p gpurun_out
(nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 1000 > gpurun_out/l2_sustained_clocks.csv &) 
timeout 120 ./scripts/l2_gather_sustained > gpurun_out/l2_gather_sustained.jsonl 2> gpurun_out/l2_gather_sustained.err; echo "rc $?"
cat gpurun_out/l2_gather_sustained.jsonl; awk 'NR%3==0' gpurun_out/l2_sustained_clocks.csv | head -12 | tr '\n' ';'
