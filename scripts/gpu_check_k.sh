#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spmm.py -m gpu -q --timeout=600 -k "slq" > gpurun_out/pytest_sign.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sign.log; tail -5 gpurun_out/pytest_sign.log; grep "^E " gpurun_out/pytest_sign.log | head
KR_BENCH_EDGES=0 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_sign.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sign.log') if x.startswith('{')]
d=json.loads(l[-1]); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'e2e ms',d['e2e']['ms_per_step'],'sign',d['e2e_sign_probes'])
PY
tail -3 gpurun_out/bench_sign.log | cut -c1-300
python scripts/time_set_edges.py 2>&1 | tail -2
