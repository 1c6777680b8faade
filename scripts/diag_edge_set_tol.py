"""How far is the wide-block (rk = 20) trace_fun_update from the oracle today? (test tolerance 1e-8)"""
import sys, os, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import numpy as np
import krylov_robustness_b200 as kr
import oracle as O
from conftest import load_graph
from test_gpu_krylov import _omega
for g in ("oregon_A1", "oregon_A0", "transport_Rome"):
    A = load_graph(g)
    n = A.shape[0]
    Om, _ = _omega(A, 10, 3)
    U, B = O.edge2low_rank(Om, n)
    nrm, _ = O.normest(A, 1e-2)
    tol = 1e-6 * np.exp(nrm)
    ox, oit, _ = O.trace_fun_update(A, U.toarray(), B, tol)
    x, it, _ = kr.trace_fun_update(A, U.toarray(), B, tol)
    print(g, "rk", U.shape[1], "it", it, oit, "rel dev %.2e" % (abs(x - ox) / abs(ox)))
