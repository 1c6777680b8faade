#!/bin/bash
# ncu --set full of the FP64 tensor-core Gram kernel (block Arnoldi, bs = 64, n = 200k)
mkdir -p gpurun_out
cat > /tmp/gram_case.py <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np
import krylov_robustness_b200 as kr
from krylov_robustness_b200.graphs import power_law_graph
A = power_law_graph(200_000, 4_000_000, 2.2, seed=3)
A = (A * (1.0 / 64.0)).tocsr()
b = np.random.default_rng(0).standard_normal((A.shape[0], 64))
M = kr.Matrix(A)
t0 = time.perf_counter()
V, K, H, p, l = kr.arnoldi_krylov(M, b)
for _ in range(4):
    V, K, H, p, l = kr.arnoldi_krylov(V, K, H, p)
print("arnoldi bs=64, 5 steps:", time.perf_counter() - t0, "s; orth", np.linalg.norm(V.T @ V - np.eye(V.shape[1])))
PY
python /tmp/gram_case.py > gpurun_out/gram_case.log 2>&1; tail -2 gpurun_out/gram_case.log
ncu --set full --clock-control none --import-source on -k regex:gram_dmma_kernel -s 6 -c 1 -o gpurun_out/prof_gram python /tmp/gram_case.py > gpurun_out/ncu_gram.log 2>&1
ls -la gpurun_out/prof_gram.ncu-rep
