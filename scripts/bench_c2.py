#!/usr/bin/env python
"""Config C2 (BASELINE.json configs[1]): weighted trace(sinh(A)) objective with the gradient over ALL edges of a
road network from datasets_paper/Transport, on one B200.

  python scripts/bench_c2.py [--graph transport_Vermont] [--check 24]

A <- A / max(A) (Tests/test_weighted_sinh_lbfgs.m:52), Omega = all lower-triangular edges,
X = 0.1 * A_Omega * uniform(0,1) (seed 4), f = sinh, df = cosh, tol = 1e-8 * cosh(normest(A)).
The gradient is  gr = -2 cosh(A + Delta(X))_Omega  (functions.fun_and_grad_all_edges: one single-vector Arnoldi
space per distinct row index, all advanced by one wide SpMM per step, chunked to fit HBM); the objective change is
the difference of two SLQ traces.  --check K compares K of the gradient entries with the oracle's
function_multiple_entries (host, one entry at a time).  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph", default="transport_Vermont")
    ap.add_argument("--check", type=int, default=24)
    ap.add_argument("--probes", type=int, default=64)
    args = ap.parse_args()
    warnings.simplefilter("ignore")
    import krylov_robustness_b200 as kr
    from conftest import load_graph
    A = load_graph(args.graph).astype(np.float64)
    A = (A / A.max()).tocsr()
    n = A.shape[0]
    L = sp.tril(A, -1).tocoo()
    Om = np.stack([L.row + 1, L.col + 1], 1).astype(np.int64)
    X = 0.1 * L.data * np.random.default_rng(4).random(L.nnz)
    ctx = kr.Context.default()
    nrm = float(kr.normest(A, 1e-2)[0])
    tol = 1e-8 * float(np.cosh(nrm))
    kr.fun_and_grad_all_edges(X[:64], A, Om[:64], "sinh", "cosh", tol, 100)     # warm-up
    ctx.sync()
    c0 = ctx.counters()
    t0 = time.perf_counter()
    gr = kr.fun_and_grad_all_edges(X, A, Om, "sinh", "cosh", tol, 100)
    ctx.sync()
    t_grad = time.perf_counter() - t0
    c1 = ctx.counters()
    # objective: -(trace sinh(A + Delta) - trace sinh(A)) by SLQ with common probes
    D = sp.csr_matrix((X, (Om[:, 0] - 1, Om[:, 1] - 1)), shape=(n, n))
    At = (A + D + D.T).tocsr()
    Z = kr.rademacher_host(n, args.probes, 9)
    t0 = time.perf_counter()
    f = -(kr.slq_trace(At, Z, 40, "sinh") - kr.slq_trace(A, Z, 40, "sinh"))
    t_obj = time.perf_counter() - t0
    out = {"workload": "C2: %s, gradient of trace sinh(A+X) over all %d edges" % (args.graph, Om.shape[0]),
           "n": n, "nnz": int(A.nnz), "edges": int(Om.shape[0]), "distinct_rows": int(np.unique(Om[:, 0]).size),
           "gradient_s": t_grad, "gradient_entries_per_s": Om.shape[0] / t_grad, "matvecs": c1["matvecs"] - c0["matvecs"],
           "launches": c1["launches"] - c0["launches"], "objective_slq_s": t_obj, "objective": f, "tol": tol}
    # where the gradient's time goes: host-side assembly of A + Delta, matrix analysis + upload, the device call
    t0 = time.perf_counter()
    Mt = kr.Matrix(At, ctx)
    ctx.sync()
    out["matrix_create_s"] = time.perf_counter() - t0
    calls = []
    for _ in range(3):
        t0 = time.perf_counter()
        v, _ = kr.function_multiple_entries(Mt, Om, "cosh", tol, 100)
        ctx.sync()
        calls.append(time.perf_counter() - t0)
    out["entries_call_s_prebuilt_matrix"] = min(calls)
    out["entries_equal_to_gradient_call"] = bool(np.array_equal(-2.0 * v, gr))
    if os.environ.get("KR_ENTRIES_LOCAL") != "0" and args.graph == "transport_Vermont" and not os.environ.get("KR_C2_SKIP_DENSE"):
        os.environ["KR_ENTRIES_LOCAL"] = "0"
        t0 = time.perf_counter()
        vd, _ = kr.function_multiple_entries(Mt, Om, "cosh", tol, 100)
        ctx.sync()
        out["entries_call_s_dense_batch"] = time.perf_counter() - t0
        out["max_abs_local_minus_dense"] = float(np.max(np.abs(v - vd)))
        del os.environ["KR_ENTRIES_LOCAL"]
    if args.check:
        import oracle as O
        idx = np.linspace(0, Om.shape[0] - 1, args.check).astype(int)
        t0 = time.perf_counter()
        ov, _ = O.function_multiple_entries(At, Om[idx], "cosh", tol, 100)
        t_or = time.perf_counter() - t0
        ref = -2.0 * ov
        out["oracle_check_entries"] = int(idx.size)
        out["oracle_s_per_entry"] = t_or / idx.size
        # cosh has even powers only and road networks are almost bipartite locally: most entries over EDGES are
        # tiny or exactly zero, so the error is measured against the largest entry (and against tol)
        out["max_abs_err_vs_oracle"] = float(np.max(np.abs(gr[idx] - ref)))
        out["max_abs_gradient_entry"] = float(np.max(np.abs(gr)))
        out["max_err_over_max_entry"] = float(np.max(np.abs(gr[idx] - ref)) / max(np.max(np.abs(ref)), 1e-300))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
