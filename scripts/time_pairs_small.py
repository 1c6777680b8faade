#!/usr/bin/env python
"""Single-launch candidate scoring (pairs_small.cuh) against the batched slot pipeline (pairs.cuh, forced with
KR_PAIR_SLOTS) on the reference's small graphs: wall time per call for 1 / 8 / 148 / 250 candidates."""
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


def best(f, reps=5):
    f()
    b = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        b = min(b, time.perf_counter() - t0)
    return b * 1e3


def main():
    import krylov_robustness_b200 as kr
    import oracle as O
    from conftest import load_graph
    for name in sys.argv[1:] or ["oregon_A8", "transport_Rome", "oregon_A7"]:
        A = load_graph(name)
        n = A.shape[0]
        M = kr.Matrix(A)
        c = O.compute_centrality(A, "eig")
        nrm = float(O.normest(A, 1e-2)[0])
        tol = 1e-6 * float(np.exp(nrm))
        E = O.find_top_edges(A, c, 250, "min")
        out = {"graph": name, "n": n, "nnz": int(A.nnz)}
        for npairs in (1, 8, 148, 250):
            Es = E[:npairs]
            os.environ.pop("KR_PAIR_SLOTS", None)
            x, it, _ = kr.trace_fun_update_edges(M, Es, -1.0, tol, 100, "exp")
            ms_small = best(lambda: kr.trace_fun_update_edges(M, Es, -1.0, tol, 100, "exp"))
            os.environ["KR_PAIR_SLOTS"] = "4096"
            ms_batched = best(lambda: kr.trace_fun_update_edges(M, Es, -1.0, tol, 100, "exp"))
            os.environ.pop("KR_PAIR_SLOTS", None)
            out["np_%d" % npairs] = {"single_launch_ms": ms_small, "batched_ms": ms_batched, "steps_mean": float(it.mean()),
                                     "steps_max": int(it.max())}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
