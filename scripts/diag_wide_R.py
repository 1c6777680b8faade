#!/usr/bin/env python
"""Diagonal of the R factors (sub-diagonal blocks of H) of the wide-block Lanczos on a scaled reference graph,
device vs oracle: are there exactly-zero / noise-level pivots (rank-deficient blocks)?"""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import krylov_robustness_b200 as kr
import oracle as O
from conftest import load_graph
name = sys.argv[1] if len(sys.argv) > 1 else "grid_England"
dfun = sys.argv[2] if len(sys.argv) > 2 else "sinh"
A = load_graph(name); A = (A / A.max()).tocsr(); n = A.shape[0]
nrm, _ = O.normest(A, 1e-2)
c = O.compute_centrality(A, "eig")
E = O.find_top_edges(A, c, 100, "min")
f_ = {"sinh": np.sinh, "cosh": np.cosh}
vals, _ = O.function_multiple_entries(A, E, dfun, 1e-6 * float(f_[dfun](nrm)), 100)
Om = E[np.argsort(-vals, kind="stable")[:30]]
aux = np.unique(Om.ravel()); k = len(aux)
U = np.zeros((n, k)); U[aux - 1, np.arange(k)] = 1
V, H, p, _ = kr.lanczos_krylov(A, U)
oV, oH, op_, _ = O.lanczos_krylov(A, U)
for j in range(1, 6):
    R = H[j * k:(j + 1) * k, (j - 1) * k:j * k]; oR = oH[j * k:(j + 1) * k, (j - 1) * k:j * k]
    d, od = np.abs(np.diag(R)), np.abs(np.diag(oR))
    print("step", j, "min|diag R| device %.3e oracle %.3e | #<1e-8: %d / %d | max rel diff of |diag| %.2e | ||V'V-I|| %.1e" % (
        d.min(), od.min(), (d < 1e-8).sum(), (od < 1e-8).sum(), np.max(np.abs(d - od) / np.maximum(od, 1e-300)),
        np.linalg.norm(V.T @ V - np.eye(V.shape[1]))))
    small = np.where(od < 1e-6)[0]
    if small.size:
        print("   small pivots (oracle):", [(int(i), float(od[i]), float(d[i])) for i in small[:6]])
    V, H, p, _ = kr.lanczos_krylov(V, H, p)
    oV, oH, op_, _ = O.lanczos_krylov(oV, oH, op_)
