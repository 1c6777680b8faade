#!/usr/bin/env python
"""Replay of the reference's unweighted break/make experiment (Tests/test_unweighted_break.m:42-95,
Tests/test_unweighted_make.m) on the device engine - row f4 of SURVEY.md section 8, config C1.

  python scripts/replay_unweighted.py [--graphs oregon_A0,oregon_A1] [--k 50] [--Q 250] [--oracle]

For each graph: symmetrise / drop self loops / largest component is already applied to the fixtures
(scripts/make_fixtures.py); nrm = exp(normest(A,1e-2)) (:56); eigenvector centrality (:63);
GREEDY_KRYLOV_BREAK with the 'min' ordering (:74) and the EIGENV heuristic (:110-125); trace(exp(A)) by
stochastic Lanczos quadrature for the relative variation.  --oracle additionally times the NumPy/SciPy
restatement of the same greedy loop on the host (the reference itself is MATLAB-only).
Prints one JSON line per (graph, method, miobi).
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
    ap.add_argument("--graphs", default="oregon_A0,oregon_A1,oregon_A8,transport_Rome")
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--Q", type=int, default=250)
    ap.add_argument("--oracle", action="store_true")
    args = ap.parse_args()
    warnings.simplefilter("ignore")
    import krylov_robustness_b200 as kr
    from conftest import load_graph
    for name in args.graphs.split(","):
        A = load_graph(name)
        n, m = A.shape[0], A.nnz // 2
        M = kr.Matrix(A)
        nrm = float(np.exp(kr.normest(M, 1e-2)[0]))
        tol = 1e-6 * nrm
        Z = kr.rademacher_host(n, 256, 7)
        trexp = kr.slq_trace(M, Z, min(60, n), "exp")
        c = kr.compute_centrality(M, "eig")
        for miobi in ("break", "make"):
            k = args.k
            Q = int(min(m - k, args.Q)) if miobi == "break" else args.Q
            kr.Context.default().sync()
            t0 = time.perf_counter()
            edges, dtr, A_new = kr.greedy_krylov(A, k, Q, c, "min", tol, 100, np.inf, 0, miobi)
            t_dev = time.perf_counter() - t0
            line = {"graph": name, "n": n, "m": m, "method": "GREEDY_KRYLOV_" + miobi.upper(), "k": k, "Q": Q,
                    "searchspace": Q + k, "time_s": t_dev, "tr_variation": dtr / trexp, "impl": "b200",
                    "edges_per_s": k * Q / t_dev}
            if args.oracle:
                import oracle as O
                t0 = time.perf_counter()
                oe, odtr, _ = O.greedy_krylov(A, k, Q, c, "min", tol, 100, np.inf, 0, miobi)
                line.update({"oracle_time_s": time.perf_counter() - t0, "same_edges": bool(np.array_equal(oe, edges)),
                             "oracle_rel_diff": abs(odtr - dtr) / abs(odtr)})
            print(json.dumps(line), flush=True)
        # EIGENV heuristic (Tests/test_unweighted_break.m:110-125)
        ind = np.argsort(-c, kind="stable")[:int(np.ceil(n / 5))]
        As = A[ind][:, ind]
        if As.nnz < 2 * args.k:
            As, ind = A, np.arange(n)
        EE = kr.find_top_edges(As, c[ind], args.k, "mult")
        H = np.stack([ind[EE[:, 0] - 1] + 1, ind[EE[:, 1] - 1] + 1], 1)
        U, B = kr.edge2low_rank(H, n)
        t0 = time.perf_counter()
        dtr, it, _ = kr.trace_fun_update(M, U.toarray(), B, tol)
        print(json.dumps({"graph": name, "method": "EIGENV", "k": args.k, "time_s": time.perf_counter() - t0,
                          "tr_variation": dtr / trexp, "block_lanczos_steps": it, "impl": "b200"}), flush=True)


if __name__ == "__main__":
    main()
