import os, sys, json, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import oracle as O, krylov_robustness_b200 as kr
from conftest import load_graph
G = json.load(open(os.path.join(ROOT, "tests/golden/oracle_golden.json")))
g = G['grad_exp/oregon_A1']; A = load_graph('oregon_A1'); n = A.shape[0]
Om = np.array(g['Omega']); X = np.array(g['X'])
U, B = O.updates._low_rank_from_omega(X, Om, n)
nrm, _ = O.normest(A, 1e-2); tol = 1e-8 * np.exp(nrm)
print("rk", U.shape[1])
for it in (1, 2, 3, 4, 6, 9, 12):
    oX, oi, _, oU = O.fun_update(A, U, B, "exp", tol, it, 0, want_basis=True)
    dX, di, _, dU = kr.fun_update(A, U, B, "exp", tol, it, 0, nargout=4)
    print("it", it, "iters", oi, di, "trace o/d", np.trace(oX), np.trace(dX), "rel", abs(np.trace(oX) - np.trace(dX)) / abs(np.trace(oX)),
          "basis orth", np.linalg.norm(dU.T @ dU - np.eye(dU.shape[1])))
# basis builders with bs = 14
b = U.copy()
V, K, H, p, l = kr.arnoldi_krylov(A, b); oV, oK, oH, op_, _ = O.arnoldi_krylov(A, b)
for s in range(4):
    print("step", s + 1, "H diff", np.abs(np.abs(H) - np.abs(oH)).max(), "|H|", np.abs(oH).max(), "orth", np.linalg.norm(V.T @ V - np.eye(V.shape[1])),
          "rel res", np.linalg.norm(A @ (V @ K) - V @ H) / np.linalg.norm(H))
    V, K, H, p, l = kr.arnoldi_krylov(V, K, H, p); oV, oK, oH, op_, _ = O.arnoldi_krylov(oV, oK, oH, op_)
