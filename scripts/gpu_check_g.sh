#!/bin/bash
mkdir -p gpurun_out
for v in degree rcm; do
  KR_BENCH_EDGES=0 KR_BENCH_REORDER=$v python bench.py --steps 2 --warmup 3 > gpurun_out/bench_t_$v.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_t_$v.log') if x.startswith('{')]
d=json.loads(l[-1]); print('$v value',d['value'],'ms/step',d['ms_per_step'],'spmm ms',d['roofline']['ms_per_launch'],'tr',d['trace_estimate'])
PY
done
