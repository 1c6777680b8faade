#!/bin/bash
mkdir -p gpurun_out
for v in pf1 pf0 pf1b; do
  lib=krylov_robustness_b200/libkrylov_b200.so
  [ $v == pf0 ] && lib=krylov_robustness_b200/libkrylov_b200_pf0.so
  KR_BENCH_EDGES=0 KR_B200_LIB=$PWD/$lib python bench.py --steps 3 --warmup 3 > gpurun_out/bench_k_$v.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_k_$v.log') if x.startswith('{')]
d=json.loads(l[-1]); print('$v value',d['value'],'ms/step',d['ms_per_step'],'spmm ms',d['roofline']['ms_per_launch'],'frac',d['roofline']['frac'],'clk',d['clocks']['sm_mhz'],'e2e',d['e2e']['value'])
PY
done
