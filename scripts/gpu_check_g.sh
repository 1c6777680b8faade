#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spmm.py -m gpu -q -x -k "not full_size" > gpurun_out/pytest_v.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_v.log
KR_BENCH_EDGES=0 timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_v.log') if x.startswith('{')]
d=json.loads(l[-1]); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'e2e ms',d['e2e']['ms_per_step'],'tr',d['trace_estimate'],d['e2e']['trace_estimate'])
PY
