#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./scripts/tma_gather > gpurun_out/tma_gather.jsonl 2> gpurun_out/tma_gather.err; echo "tma rc $?"; cat gpurun_out/tma_gather.jsonl
timeout 600 python scripts/exp_spmm.py base=libkrylov_b200.so pw8=libkrylov_b200_pw8.so > gpurun_out/exp_i.jsonl 2> gpurun_out/exp_i.err
cat gpurun_out/exp_i.jsonl; tail -3 gpurun_out/exp_i.err
timeout 900 python scripts/replay_weighted.py --graphs grid_England --oracle > gpurun_out/replay_weighted.jsonl 2> gpurun_out/replay_weighted.err
cut -c1-400 gpurun_out/replay_weighted.jsonl; tail -3 gpurun_out/replay_weighted.err
