#!/bin/bash
mkdir -p gpurun_out
KR_SPMM_DIRECT=1 timeout 180 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_expmv.py -m gpu -q -x -k "not full_size" > gpurun_out/pytest_u.log 2>&1; echo "exit $?"; tail -2 gpurun_out/pytest_u.log
for v in 1 0; do
  KR_BENCH_EDGES=0 KR_SPMM_DIRECT=$v timeout 240 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_u_$v.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_u_$v.log') if x.startswith('{')]
d=json.loads(l[-1]) if l else None
print('direct=$v', (d['value'], d['ms_per_step'], d['roofline']['ms_per_launch']) if d else 'no output')
PY
done
