#!/bin/bash
mkdir -p gpurun_out
for cap in 1024 2048 2048r256; do
  lib=krylov_robustness_b200/libkrylov_b200.so
  [ $cap != 4096 ] && lib=krylov_robustness_b200/libkrylov_b200_cap$cap.so
  KR_B200_LIB=$PWD/$lib python bench.py --steps 2 --warmup 3 > gpurun_out/bench_f_cap$cap.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_f_cap$cap.log') if x.startswith('{')]
d=json.loads(l[-1]); print('cap=$cap value',d['value'],'ms/step',d['ms_per_step'],'spmm ms',d['roofline']['ms_per_launch'],'frac',d['roofline']['frac'])
PY
done
