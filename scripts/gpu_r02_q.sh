#!/bin/bash
# round 2, call Q: small-graph latencies with the final build, greedy replay on the reference's graphs
mkdir -p gpurun_out
timeout 900 python scripts/time_small.py grid_England transport_Rome oregon_A8 > gpurun_out/r02q_time_small.jsonl 2> gpurun_out/r02q_time_small.err; echo "rc $?"; cat gpurun_out/r02q_time_small.jsonl; tail -5 gpurun_out/r02q_time_small.err
timeout 900 python scripts/replay_unweighted.py > gpurun_out/r02q_replay_unweighted.jsonl 2>&1; tail -12 gpurun_out/r02q_replay_unweighted.jsonl | cut -c1-400
KR_PAIR_EIG_JACOBI=1 timeout 600 python scripts/time_pairs_small.py oregon_A8 > gpurun_out/r02q_time_pairs_small_jacobi_ab.jsonl 2>&1; cat gpurun_out/r02q_time_pairs_small_jacobi_ab.jsonl
