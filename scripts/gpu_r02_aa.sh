#!/bin/bash
# round 2, call AA: ncu --set full of the new kernels: ts_gram at C5 scale (n = 1M, 528 x 528 node columns), node_pair_kernel,
# pair_small_kernel (250 candidates on Oregon A8)
mkdir -p gpurun_out
export KR_BENCH_C5_CAND=100000
ncu --set full --clock-control none --import-source on -k regex:"ts_gram_kernel" -s 12 -c 1 -o gpurun_out/r02aa_prof_ts_gram_c5 python scripts/bench_screen.py > gpurun_out/r02aa_ncu_ts_gram.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"node_pair_kernel" -s 1 -c 1 -o gpurun_out/r02aa_prof_node_pair python scripts/bench_screen.py > gpurun_out/r02aa_ncu_node_pair.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pair_small_kernel" -s 30 -c 1 -o gpurun_out/r02aa_prof_pair_small python scripts/time_pairs_small.py oregon_A8 > gpurun_out/r02aa_ncu_pair_small.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02aa_launches_screen_c5.csv python scripts/bench_screen.py > /dev/null 2>&1
ls -la gpurun_out | grep r02aa
