"""Candidate scoring on mid-size graphs (n = 6 k .. 100 k): which pipeline runs, time per 250-candidate round."""
import sys, time, json, warnings
import numpy as np
warnings.simplefilter("ignore")
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_graph
import oracle as O
import krylov_robustness_b200 as kr
for name in ("misc_as_735", "oregon_A7", "transport_Vermont"):
    A = load_graph(name); n = A.shape[0]
    nrm, _ = O.normest(A, 1e-2); tol = 1e-6 * float(np.exp(nrm))
    c = kr.compute_centrality(A, "eig")
    E = O.find_top_edges(A, c, 250, "min")
    M = kr.Matrix(A)
    x, it, lk = kr.trace_fun_update_edges(M, E, -1.0, tol, 100, "exp")          # warm-up
    ts = []
    for _ in range(5):
        t = time.perf_counter(); x, it, lk = kr.trace_fun_update_edges(M, E, -1.0, tol, 100, "exp"); ts.append(time.perf_counter() - t)
    t = time.perf_counter()
    for i, j in E[:20]:
        U = np.zeros((n, 2)); U[i - 1, 0] = U[j - 1, 1] = 1.0
        O.trace_fun_update(A, U, -np.array([[0, 1.0], [1.0, 0]]), tol, 100)
    t_or = (time.perf_counter() - t) / 20 * 250
    print(json.dumps({"graph": name, "n": n, "nnz": int(A.nnz), "candidates": 250, "mean_steps": float(np.mean(it)),
                      "device_ms_per_round": round(1e3 * min(ts), 2), "oracle_ms_per_round_est": round(1e3 * t_or, 1)}))
