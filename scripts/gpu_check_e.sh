#!/bin/bash
mkdir -p gpurun_out
KR_SPMM_UNROLL=8 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_e.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_e.log
tail -5 gpurun_out/pytest_e.log
for U in 4 8; do
  KR_SPMM_UNROLL=$U python bench.py --steps 2 --warmup 3 > gpurun_out/bench_e_u$U.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_e_u$U.log') if x.startswith('{')]
d=json.loads(l[-1]); print('U=$U value',d['value'],'ms/step',d['ms_per_step'],'spmm ms',d['roofline']['ms_per_launch'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
PY
done
