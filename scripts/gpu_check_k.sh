#!/bin/bash
mkdir -p gpurun_out
for v in pool48 pool96; do
  lib=krylov_robustness_b200/libkrylov_b200.so
  [ $v == pool48 ] && lib=krylov_robustness_b200/libkrylov_b200_pool48.so
  KR_B200_LIB=$PWD/$lib timeout 600 python scripts/bench_edges.py --ncand 6000 > gpurun_out/bench_edges_$v.json 2> gpurun_out/bench_edges_$v.err; echo "$v rc $?"; cut -c1-330 gpurun_out/bench_edges_$v.json; tail -2 gpurun_out/bench_edges_$v.err
done
