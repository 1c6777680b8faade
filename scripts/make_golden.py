"""Generate tests/golden/oracle_golden.json: outputs of the oracle (oracle/) on seeded inputs built from the
committed graph fixtures.  The reference ships no golden vectors; these pin the ORACLE against
drift and give the GPU tests fixed numbers to hit without re-running the oracle.

  python scripts/make_golden.py        (CPU only, ~1 minute)
"""
import json
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import oracle as O  # noqa: E402
from conftest import load_graph, edge_UB  # noqa: E402


def omega(A, k, seed, min_degree=3):
    L = sp.tril(A, -1).tocoo()
    deg = np.diff(sp.csr_matrix(A).indptr)
    ok = np.where((deg[L.row] >= min_degree) & (deg[L.col] >= min_degree))[0]
    rng = np.random.default_rng(seed)
    sel = ok[rng.choice(ok.size, k, replace=False)]
    return np.stack([L.row[sel] + 1, L.col[sel] + 1], 1), 0.1 * L.data[sel] * rng.random(k)


def main():
    G = {}
    # --- candidate-edge trace updates (krylov_miobi's loop)
    for gname, fun, sign in (("oregon_A0", "exp", -1.0), ("oregon_A8", "exp", 1.0), ("transport_Rome", "sinh", -1.0)):
        A = load_graph(gname)
        n = A.shape[0]
        f = {"exp": np.exp, "sinh": np.sinh, "cosh": np.cosh}[fun]
        nrm, _ = O.normest(A, 1e-2)
        tol = 1e-6 * abs(float(f(nrm)))
        c = O.compute_centrality(A, "eig")
        E = O.find_top_edges(A, c, 16, "min") if sign < 0 else O.find_top_missing_edges(A, c, 16, "min")
        xs, its = [], []
        for i, j in E:
            U, B = edge_UB(n, int(i), int(j), sign)
            x, it, _ = O.trace_fun_update(A, U, B, tol, 100, 0, fun)
            xs.append(x)
            its.append(it)
        G["edges/%s/%s/%+d" % (gname, fun, int(sign))] = {"E": E.tolist(), "tol": tol, "Xm": xs, "iter": its}
    # --- expmv counters and a checksum of the result
    for gname, q in (("oregon_A0", 10), ("transport_Rome", 16), ("grid_Mexico", 1)):
        A = load_graph(gname)
        b = np.sign(np.random.default_rng(q).standard_normal((A.shape[0], q)))
        f, s, m, mv, mvd, unA = O.expmv(1, A, b)
        G["expmv/%s/%d" % (gname, q)] = {"s": int(s), "m": int(m), "mv": int(mv), "mvd": int(mvd), "unA": int(unA),
                                         "colsum": f.sum(axis=0).tolist(), "fro": float(np.linalg.norm(f))}
    # --- entries of f(A)
    A = load_graph("oregon_A0")
    n = A.shape[0]
    rng = np.random.default_rng(5)
    om = np.stack([rng.integers(1, n + 1, 24), rng.integers(1, n + 1, 24)], 1)
    om[5:9, 0] = om[5, 0]
    nrm, _ = O.normest(A, 1e-2)
    X, it = O.function_multiple_entries(A, om, "exp", 1e-8 * np.exp(nrm), 100)
    G["entries/oregon_A0/exp"] = {"omega": om.tolist(), "tol": 1e-8 * float(np.exp(nrm)), "X": X.tolist(), "iter": int(it)}
    # --- objective + gradient callbacks and Hessian
    A = load_graph("oregon_A1")
    Om, Xw = omega(A, 12, 5)   # 8 edges hit a rank-deficient residual block
    import scipy.linalg as sla
    eA = sla.expm(A.toarray())
    eAo = [float(eA[a - 1, b - 1]) for a, b in Om]
    f, gr = O.fun_and_grad_krylov_exp(Xw, A, Om, np.array(eAo), 1e-8, 100)
    G["grad_exp/oregon_A1"] = {"Omega": Om.tolist(), "X": Xw.tolist(), "eA": eAo, "f": f, "gr": gr.tolist()}
    A = load_graph("grid_England")
    Om, Xw = omega(A, 6, 2, min_degree=2)
    H = O.hessianfcn_exp(Xw, A, Om, 1e-10, 60)
    G["hessian_exp/grid_England"] = {"Omega": Om.tolist(), "X": Xw.tolist(), "Hes": H.tolist()}
    # --- SLQ (throughput mode) with the counter-based probes
    from krylov_robustness_b200.engine import rademacher_host
    A = (load_graph("oregon_A8") / 8.0).tocsr()
    Z = rademacher_host(A.shape[0], 24, 7)
    tr, vals, al, be = O.slq_trace(A, Z, 20, "exp")
    G["slq/oregon_A8_div8/exp"] = {"seed": 7, "k": 24, "m": 20, "tr": tr, "vals": vals.tolist()}
    # --- greedy drivers
    A = load_graph("transport_Barcelona")
    nrm, _ = O.normest(A, 1e-2)
    c = O.compute_centrality(A, "eig")
    for miobi in ("break", "make"):
        e, rob, _ = O.greedy_krylov(A, 4, 30, c, "min", 1e-6 * np.exp(nrm), 100, np.inf, 0, miobi)
        G["greedy/transport_Barcelona/%s" % miobi] = {"centrality": c.tolist(), "tol": 1e-6 * float(np.exp(nrm)),
                                                       "edges": e.tolist(), "rob": rob}
    out = os.path.join(ROOT, "tests", "golden", "oracle_golden.json")
    with open(out, "w") as fh:
        json.dump(G, fh)
    print("wrote", out, os.path.getsize(out), "bytes,", len(G), "cases")


if __name__ == "__main__":
    main()
