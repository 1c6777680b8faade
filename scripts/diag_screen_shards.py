"""Where does a screened round spend its time per candidate shard? (diagnostic for the N = 2 bench)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krylov_robustness_b200 as kr  # noqa: E402
from krylov_robustness_b200.graphs import power_law_graph, spectral_radius_estimate  # noqa: E402

A = power_law_graph(1_000_000, 20_000_000, 2.2, 20260310)
lam = spectral_radius_estimate(A, 30)
A = (A * (1.0 / lam)).tocsr()
M = kr.Matrix(A)
cvec = kr.compute_centrality(M, "eig", 1e-10)
E = kr.find_top_missing_edges(A, cvec, 100_000, "min")
tol = 1e-6 * float(np.e)
kr.greedy_round(M, E[:4000], 1.0 / lam, tol, 100, "exp", "make", screen=True)
for name, sl in (("all", slice(0, 100000)), ("first half", slice(0, 50000)), ("second half", slice(50000, 100000)),
                 ("second half again", slice(50000, 100000)), ("last eighth", slice(87500, 100000))):
    c0 = M.ctx.counters()
    t0 = time.perf_counter()
    b, v, sc, mask, info = kr.greedy_round(M, E[sl], 1.0 / lam, tol, 100, "exp", "make", screen=True)
    dt = time.perf_counter() - t0
    c1 = M.ctx.counters()
    print(json.dumps({"shard": name, "seconds": dt, "info": info, "launches": c1["launches"] - c0["launches"],
                      "matvecs": c1["matvecs"] - c0["matvecs"]}), flush=True)
