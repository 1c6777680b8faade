#!/bin/bash
mkdir -p gpurun_out
run() { # name lib unroll
  KR_SPMM_UNROLL=$3 KR_B200_LIB=$PWD/krylov_robustness_b200/$2 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_j_$1.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_j_$1.log') if x.startswith('{')]
d=json.loads(l[-1]); print('$1 value',d['value'],'ms/step',d['ms_per_step'],'spmm ms',d['roofline']['ms_per_launch'],'frac',d['roofline']['frac'])
PY
}
run cap4096_u8 libkrylov_b200.so 8
run cap4096_u4 libkrylov_b200.so 4
run cap2048_u8 libkrylov_b200_cap2048.so 8
run cap6144_u8 libkrylov_b200_cap6144.so 8
run cap6144_u4 libkrylov_b200_cap6144.so 4
