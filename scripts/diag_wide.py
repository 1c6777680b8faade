#!/usr/bin/env python
"""Per-step comparison of the wide-block trace_fun_update (device vs oracle) on a scaled reference graph:
calling with it = j returns the value after exactly j steps."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import krylov_robustness_b200 as kr
import oracle as O
from conftest import load_graph

name = sys.argv[1] if len(sys.argv) > 1 else "grid_England"
fun, dfun = (sys.argv[2], sys.argv[3]) if len(sys.argv) > 3 else ("cosh", "sinh")
A = load_graph(name); A = (A / A.max()).tocsr(); n = A.shape[0]
nrm, _ = O.normest(A, 1e-2)
c = O.compute_centrality(A, "eig")
E = O.find_top_edges(A, c, 100, "min")
f_ = {"sinh": np.sinh, "cosh": np.cosh}
vals, _ = O.function_multiple_entries(A, E, dfun, 1e-6 * float(f_[dfun](nrm)), 100)
ind = np.argsort(-vals, kind="stable")[:30]
Om = E[ind]; x = 0.05 * np.ones(30)
aux = np.unique(Om.ravel()); k = len(aux); pos = {a: i for i, a in enumerate(aux)}
U = np.zeros((n, k)); U[aux - 1, np.arange(k)] = 1
B = np.zeros((k, k))
for (a, b), v in zip(Om, x): B[pos[a], pos[b]] = v; B[pos[b], pos[a]] = v
M = kr.Matrix(A)
tol = 1e-6 * float(f_[fun](nrm)) * float(f_[fun](nrm))
print("graph", name, "n", n, "k", k, "tol_eff", tol)
for j in range(1, 12):
    xd, itd, ld = kr.trace_fun_update(M, U, B, 0.0, j, 0, fun)
    xo, ito, lo = O.trace_fun_update(A, U, B, 0.0, j, 0, fun)
    print(j, "device %.15e it %d lucky %d | oracle %.15e it %d lucky %d | rel diff %.2e" % (xd, itd, ld, xo, ito, lo, abs(xd - xo) / abs(xo)))
xd, itd, _ = kr.trace_fun_update(M, U, B, tol, 100, 0, fun)
xo, ito, _ = O.trace_fun_update(A, U, B, tol, 100, 0, fun)
print("with tol: device", xd, itd, "oracle", xo, ito)
