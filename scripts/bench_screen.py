#!/usr/bin/env python
"""Config C5 as a greedy 'make' round: 10^5 candidates of find_top_missing_edges on the C3 graph through
kr_greedy_round with the node-basis screen (csrc/nodepairs.cuh); the exact round (all candidates through the block
Lanczos pipeline) for comparison when KR_SCREEN_EXACT=1 (17 s)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import krylov_robustness_b200 as kr
    from krylov_robustness_b200.graphs import power_law_graph, spectral_radius_estimate
    n = int(os.environ.get("KR_BENCH_N", 1_000_000))
    nnz = int(os.environ.get("KR_BENCH_NNZ", 20_000_000))
    ncand = int(os.environ.get("KR_BENCH_C5_CAND", 100_000))
    A = power_law_graph(n, nnz, 2.2, 20260310)
    lam = spectral_radius_estimate(A, 30)
    A = (A * (1.0 / lam)).tocsr()
    M = kr.Matrix(A)
    cvec = kr.compute_centrality(M, "eig", 1e-10)
    E = kr.find_top_missing_edges(A, cvec, ncand, "min")
    tol = 1e-6 * float(np.e)
    kr.greedy_round(M, E[:4000], 1.0 / lam, tol, 100, "exp", "make", screen=True)       # warm-up
    c0 = M.ctx.counters()
    t0 = time.perf_counter()
    b, v, scores, mask, info = kr.greedy_round(M, E, 1.0 / lam, tol, 100, "exp", "make", screen=True)
    dt = time.perf_counter() - t0
    c1 = M.ctx.counters()
    out = {"candidates": int(E.shape[0]), "seconds_screened_round": dt, "edges_per_sec": E.shape[0] / dt, "info": info,
           "best_candidate": [int(x) for x in E[b]], "best_value": v, "launches": c1["launches"] - c0["launches"],
           "matvecs": c1["matvecs"] - c0["matvecs"]}
    if os.environ.get("KR_SCREEN_EXACT", "0") not in ("", "0"):
        t0 = time.perf_counter()
        x, it, _ = kr.trace_fun_update_edges(M, E, 1.0 / lam, tol, 100, "exp")
        dte = time.perf_counter() - t0
        be, ve = kr.select_candidate(x, "make")
        rel = np.abs(scores - x) / np.abs(x)
        out.update({"seconds_exact_round": dte, "edges_per_sec_exact": E.shape[0] / dte, "same_winner": bool(be == b and ve == v),
                    "screen_max_rel_dev": float(rel.max()), "screen_median_rel_dev": float(np.median(rel[~mask])),
                    "speedup": dte / dt})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
