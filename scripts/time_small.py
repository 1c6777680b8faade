#!/usr/bin/env python
"""Where does the time go on the reference's own (small) graphs?  Times every host-API call of the weighted
experiment (scripts/replay_weighted.py) on the device engine and on the oracle, call by call."""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")


def t(f, reps=3):
    f()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def main():
    import krylov_robustness_b200 as kr
    import oracle as O
    from conftest import load_graph
    for name in sys.argv[1:] or ["grid_England", "transport_Rome", "oregon_A8"]:
        A = load_graph(name)
        A = (A / A.max()).tocsr()
        n = A.shape[0]
        out = {"graph": name, "n": n}
        for tag, P in (("b200", kr), ("oracle", O)):
            Ad = kr.Matrix(A) if tag == "b200" else A
            c = P.compute_centrality(Ad if tag == "b200" else A, "eig")
            nrm = P.normest(Ad, 1e-2)
            nrm = float(nrm[0] if isinstance(nrm, tuple) else nrm)
            E = P.find_top_edges(A, c, 100, "min")
            tol_df = 1e-6 * float(np.cosh(nrm))
            vals, _ = P.function_multiple_entries(Ad, E, "cosh", tol_df, 100)
            ind = np.argsort(-vals, kind="stable")[:30]
            Om, dfA = E[ind], vals[ind]
            x = 0.05 * np.ones(30)
            tol = 1e-6 * float(np.sinh(nrm))
            r = {
                "centrality_ms": t(lambda: P.compute_centrality(Ad if tag == "b200" else A, "eig"), 1),
                "normest_ms": t(lambda: P.normest(Ad, 1e-2)),
                "entries_100pairs_ms": t(lambda: P.function_multiple_entries(Ad, E, "cosh", tol_df, 100)),
                "fun_and_grad_30edges_ms": t(lambda: P.fun_and_grad_krylov_fun(x, Ad, Om, "sinh", "cosh", dfA, tol, 100)),
                "trace_fun_update_rank2_ms": t(lambda: P.trace_fun_update(Ad, *P.edge2low_rank(Om[:1], n, 1.0)[:2], tol, 100, 0, "sinh")),
            }
            if tag == "b200":
                r["trace_fun_update_edges_250_ms"] = t(lambda: kr.trace_fun_update_edges(Ad, P.find_top_edges(A, c, 250, "min"), -1.0, tol, 100, "sinh"))
            out[tag] = r
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
