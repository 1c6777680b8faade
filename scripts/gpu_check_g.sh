#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_expmv.py -m gpu -q -x > gpurun_out/pytest_r.log 2>&1; tail -3 gpurun_out/pytest_r.log
python - > gpurun_out/trace_exp_time.log 2>&1 <<PY
import sys, time, warnings; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
warnings.simplefilter('ignore')
import numpy as np, oracle as O, krylov_robustness_b200 as kr
from conftest import load_graph
for g in ('oregon_A0','oregon_A8'):
    A=load_graph(g); n=A.shape[0]; rng=np.random.default_rng(0)
    probes=[(np.sign(rng.standard_normal((n,10))),np.sign(rng.standard_normal((n,10)))) for _ in range(34)]
    M=kr.Matrix(A); kr.trace_exp(M,probes)
    t=time.perf_counter(); tr=kr.trace_exp(M,probes); td=time.perf_counter()-t
    t=time.perf_counter(); otr=O.trace_exp(A,probes=probes); to=time.perf_counter()-t
    print(g,'device',td,'oracle',to,'rel',abs(tr-otr)/abs(otr))
PY
cat gpurun_out/trace_exp_time.log
