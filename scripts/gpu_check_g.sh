#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spmm.py -m gpu -q -x -k "not full_size" > gpurun_out/pytest_q.log 2>&1; tail -1 gpurun_out/pytest_q.log
KR_B200_LIB=$PWD/krylov_robustness_b200/libkrylov_b200_t128.so python -m pytest tests/test_gpu_spmm.py -m gpu -q -x -k "not full_size" > gpurun_out/pytest_q128.log 2>&1; tail -1 gpurun_out/pytest_q128.log
for v in default t128; do
  lib=krylov_robustness_b200/libkrylov_b200.so; [ $v = t128 ] && lib=krylov_robustness_b200/libkrylov_b200_t128.so
  KR_BENCH_EDGES=0 KR_B200_LIB=$PWD/$lib python bench.py --steps 2 --warmup 3 > gpurun_out/bench_q_$v.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_q_$v.log') if x.startswith('{')]
d=json.loads(l[-1]); print('$v value',d['value'],'ms/step',d['ms_per_step'],'spmm ms',d['roofline']['ms_per_launch'])
PY
done
