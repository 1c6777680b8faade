#!/bin/bash
# round 2, call L: TMA gather4 micro-benchmark, column-slab experiment, ncu of the 256-bit SpMM / ts_gram / launch list
mkdir -p gpurun_out
for b in 1 4; do timeout 120 scripts/tma_gather4 $b > gpurun_out/r02l_tma_gather4_box$b.jsonl 2>&1; echo "gather4 box $b rc $?"; tail -3 gpurun_out/r02l_tma_gather4_box$b.jsonl; done
timeout 600 python scripts/exp_slabs.py 2 4 > gpurun_out/r02l_slabs.jsonl 2> gpurun_out/r02l_slabs.err; echo "slabs rc $?"; cat gpurun_out/r02l_slabs.jsonl; tail -3 gpurun_out/r02l_slabs.err
export KR_BENCH_M=2 KR_BENCH_EDGES=0 KR_BENCH_C4=0
python bench.py --steps 1 --warmup 3 > gpurun_out/r02l_bench_m2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -s 6 -c 1 -o gpurun_out/r02l_prof_spmm \
    python bench.py --steps 1 --warmup 3 > gpurun_out/r02l_ncu_spmm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02l_launches_bench_k512_m2.csv \
    python bench.py --steps 1 --warmup 3 > gpurun_out/r02l_ncu_launch.log 2>&1
unset KR_BENCH_M KR_BENCH_EDGES KR_BENCH_C4
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
python /tmp/gram_case.py > gpurun_out/r02l_gram_case.log 2>&1; tail -2 gpurun_out/r02l_gram_case.log
ncu --set full --clock-control none --import-source on -k regex:"ts_gram_kernel|ts_update_kernel" -s 8 -c 2 -o gpurun_out/r02l_prof_ts python /tmp/gram_case.py > gpurun_out/r02l_ncu_ts.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02l_launches_arnoldi_bs64.csv python /tmp/gram_case.py > /dev/null 2>&1
ls -la gpurun_out | grep r02l
