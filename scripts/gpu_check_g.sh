#!/bin/bash
mkdir -p gpurun_out
KR_BENCH_EDGES=0 KR_B200_LIB=$PWD/krylov_robustness_b200/libkrylov_b200_el.so python bench.py --steps 2 --warmup 3 > gpurun_out/bench_p_el.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_p_el.log') if x.startswith('{')]
d=json.loads(l[-1]); print('evict_last value',d['value'],'ms/step',d['ms_per_step'],'spmm ms',d['roofline']['ms_per_launch'])
PY
