#!/bin/bash
mkdir -p gpurun_out
python - <<PY
import torch
p=torch.cuda.get_device_properties(0); print('L2',p.L2_cache_size/2**20,'MB')
PY
run() { # name persist_mb
  KR_BENCH_EDGES=0 KR_SPMM_L2PERSIST=$2 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n_$1.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_n_$1.log') if x.startswith('{')]
d=json.loads(l[-1]); print('$1 value',d['value'],'ms/step',d['ms_per_step'],'spmm ms',d['roofline']['ms_per_launch'],'launches',d['roofline']['launches_timed'])
PY
}
run p0 0
run p32 32
run p64 64
run p96 96
