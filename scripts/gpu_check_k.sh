#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_mex_gateway.py -m gpu -q --timeout=600 > gpurun_out/pytest_mex.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mex.log; tail -5 gpurun_out/pytest_mex.log; grep "^E " gpurun_out/pytest_mex.log | head
nvidia-smi topo -m 2>/dev/null | head -12; lscpu | grep -i "numa\|socket" | head
KR_BENCH_EDGES=0 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_numa.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_numa.log') if x.startswith('{')]
d=json.loads(l[-1]); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'e2e ms',d['e2e']['ms_per_step'],'numa',d['numa'])
PY
