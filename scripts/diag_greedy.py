"""Diagnostic: run the oracle's greedy loop and, at every round, score the same candidate list on the
device; report the first round where the choice or any score / iteration count differs."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import oracle as O
import krylov_robustness_b200 as kr
from conftest import load_graph, edge_UB

name, miobi, k, Q = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
A = load_graph(name); n = A.shape[0]
nrm, _ = O.normest(A, 1e-2); tol = 1e-6 * np.exp(nrm)
c = O.compute_centrality(A, "eig")
rnd = [0]
def scorer(Acur, E, sign, rescale, tol, it):
    rnd[0] += 1
    ov, oi = [], []
    for i, j in E:
        U, B = edge_UB(n, int(i), int(j), sign)
        x, itn, _ = O.trace_fun_update(Acur, U, B / rescale, tol, it)
        ov.append(x); oi.append(itn)
    ov = np.array(ov); oi = np.array(oi)
    dv, di, _ = kr.trace_fun_update_edges(Acur, E, sign / rescale, tol, it, "exp")
    rel = np.abs(dv - ov) / np.maximum(np.abs(ov), tol)
    bad = np.where((rel > 1e-9) | (di != oi))[0]
    ob = int(np.argmin(ov)) if sign < 0 else int(np.argmax(ov))
    db = int(np.argmin(dv)) if sign < 0 else int(np.argmax(dv))
    if bad.size or ob != db:
        deg = np.diff(Acur.indptr)
        print("round", rnd[0], "oracle best", ob, E[ob], "device best", db, E[db], "n_bad", bad.size)
        for h in bad[:8]:
            print("   cand", h, E[h], "deg", deg[E[h, 0] - 1], deg[E[h, 1] - 1], "oracle", ov[h], oi[h], "device", dv[h], di[h], "rel", rel[h])
    return ov
e, rob, _ = O.greedy_krylov(A, k, Q, c, "min", tol, 100, np.inf, 0, miobi, scorer=scorer)
print("done rounds", rnd[0], "rob", rob)
