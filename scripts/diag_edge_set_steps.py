"""Per-step values of the wide-block trace_fun_update on Oregon A0 (rk = 14), device vs oracle: where do they part?"""
import sys, os, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import numpy as np
import krylov_robustness_b200 as kr
import oracle as O
from conftest import load_graph
from test_gpu_krylov import _omega
A = load_graph("oregon_A0")
n = A.shape[0]
Om, _ = _omega(A, 10, 3)
U, B = O.edge2low_rank(Om, n)
U = U.toarray()
nrm, _ = O.normest(A, 1e-2)
tol = 1e-6 * np.exp(nrm)
print("nrm", nrm, "tol", tol, "rk", U.shape[1], "deg of nodes", np.diff(A.indptr)[np.unique(Om) - 1])
prev = {}
for j in range(1, 13):
    ox, oit, olk = O.trace_fun_update(A, U, B, 0.0, j)
    x, it, lk = kr.trace_fun_update(A, U, B, 0.0, j)
    print(j, "oracle %.15e (it %d lucky %d)  device %.15e (it %d lucky %d)  rel %.2e  |dX(j)-dX(j-2)|/tol oracle %.3e device %.3e"
          % (ox, oit, olk, x, it, lk, abs(x - ox) / abs(ox),
             abs(ox - prev.get(j - 2, (0, 0))[0]) / tol, abs(x - prev.get(j - 2, (0, 0))[1]) / tol))
    prev[j] = (ox, x)
