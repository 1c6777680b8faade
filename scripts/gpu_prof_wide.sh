#!/bin/bash
mkdir -p gpurun_out
KR_PROFILE_WIDE=1 timeout 300 python - > gpurun_out/prof_wide.log 2>&1 <<'PY'
import sys, time, warnings
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
warnings.simplefilter("ignore")
import numpy as np
import krylov_robustness_b200 as kr
from conftest import load_graph
for name in ("grid_England", "transport_Rome"):
    A = load_graph(name); A = (A / A.max()).tocsr(); n = A.shape[0]
    M = kr.Matrix(A)
    c = kr.compute_centrality(M, "eig"); nrm = float(kr.normest(M, 1e-2)[0])
    E = kr.find_top_edges(A, c, 100, "min")
    vals, _ = kr.function_multiple_entries(M, E, "cosh", 1e-6 * np.cosh(nrm), 100)
    ind = np.argsort(-vals, kind="stable")[:30]
    Om, dfA = E[ind], vals[ind]
    x = 0.05 * np.ones(30)
    tol = 1e-6 * float(np.sinh(nrm))
    kr.fun_and_grad_krylov_fun(x, M, Om, "sinh", "cosh", dfA, tol, 100)
    print("====", name, "unique nodes", np.unique(Om).size, flush=True)
    t0 = time.perf_counter()
    kr.fun_and_grad_krylov_fun(x, M, Om, "sinh", "cosh", dfA, tol, 100)
    print("total ms", (time.perf_counter() - t0) * 1e3, flush=True)
PY
cat gpurun_out/prof_wide.log
