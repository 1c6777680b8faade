#!/usr/bin/env python
"""Inputs for scripts/make_reference_goldens.m (the route to PINNED parity): the same seeded inputs the parity
tests use, written as one MAT-v5 file that GNU Octave / MATLAB can load.  Deterministic; committed as
tests/golden/reference_inputs.mat.   python scripts/make_reference_inputs.py"""
import os
import sys

import numpy as np
import scipy.io as sio
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_inputs():
    from conftest import load_graph
    import oracle as O
    out = {}
    A0 = load_graph("oregon_A0")
    n0 = A0.shape[0]
    rng = np.random.default_rng(20261018)
    c0 = O.compute_centrality(A0, "eig")
    out["A0"] = sp.csc_matrix(A0)
    out["A0_break_edges"] = O.find_top_edges(A0, c0, 8, "min").astype(np.float64)
    out["A0_make_edges"] = O.find_top_missing_edges(A0, c0, 8, "min").astype(np.float64)
    out["A0_tol"] = 1e-6 * float(np.exp(O.normest(A0, 1e-2)[0]))
    out["A0_omega"] = np.stack([rng.integers(1, n0 + 1, 10), rng.integers(1, n0 + 1, 10)], 1).astype(np.float64)
    out["A0_b"] = rng.standard_normal((n0, 3))
    out["A0_centrality"] = c0.reshape(-1, 1)
    Rome = load_graph("transport_Rome")
    cR = O.compute_centrality(Rome, "eig")
    out["Rome"] = sp.csc_matrix(Rome)
    out["Rome_edges"] = O.find_top_edges(Rome, cR, 4, "min").astype(np.float64)
    out["Rome_tol"] = 1e-6 * float(np.sinh(O.normest(Rome, 1e-2)[0]))
    Mx = load_graph("grid_Mexico")
    nM = Mx.shape[0]
    nodes = np.sort(rng.choice(nM, 6, replace=False))
    U = np.zeros((nM, 6))
    U[nodes, np.arange(6)] = 1.0
    B = rng.standard_normal((6, 6)) * 0.1
    out["Mexico"] = sp.csc_matrix(Mx)
    out["Mexico_U"] = U
    out["Mexico_B"] = 0.5 * (B + B.T)
    out["Mexico_tol"] = 1e-9
    # probes for mc_trace(A0, n, 1e-3, 60, 1): K = ceil(60/30) = 2 iterations, S then G each (mc_trace.m:43-44)
    out["A0_probes"] = np.sign(rng.standard_normal((n0, 40)))
    # ---- second batch (drawn AFTER everything above, so the first batch is unchanged)
    # trace_exp(A0) = mc_trace(Afun, n, 1e-4, 1000, 1): K = ceil(1000/30) = 34 iterations at most, 20 probe columns each
    out["A0_probes_long"] = np.sign(rng.standard_normal((n0, 680)))
    hubs = np.argsort(-c0, kind="stable")
    out["A0_self_node"] = float(hubs[2] + 1)
    out["A0_set_edges"] = O.find_top_edges(A0, c0, 40, "min")[[0, 7, 19, 33]].astype(np.float64)
    mixed = O.find_top_edges(A0, c0, 6, "mult").astype(np.float64)
    mixed[1] = [hubs[0] + 1, hubs[0] + 1]                       # a self loop among the candidates (krylov_miobi.m:88-99)
    out["A0_mixed_edges"] = mixed
    out["A0_loop_nodes"] = np.sort(rng.choice(n0, 10, replace=False) + 1).astype(np.float64).reshape(-1, 1)
    # weighted-experiment shapes on the Mexican grid (Tests/test_weighted_*: Omega = modifiable edges, X = weights)
    T = sp.tril(Mx, -1).tocoo()
    pick = np.sort(rng.choice(T.nnz, 6, replace=False))
    out["Mexico_Omega"] = np.stack([T.row[pick] + 1, T.col[pick] + 1], 1).astype(np.float64)
    out["Mexico_X"] = (0.1 * rng.uniform(0.0, 1.0, 6)).reshape(-1, 1)
    # ---- third batch: the call shape of Tests/test_weighted_*_lbfgs.m (30 modifiable edges -> a block of ~55 columns;
    # on a 552-node grid the Krylov space saturates and fun_update.m:84-90 switches to dense arithmetic)
    pick30 = np.sort(rng.choice(T.nnz, 30, replace=False))
    out["Mexico_Omega30"] = np.stack([T.row[pick30] + 1, T.col[pick30] + 1], 1).astype(np.float64)
    out["Mexico_X30"] = (0.1 * rng.uniform(0.0, 1.0, 30)).reshape(-1, 1)
    # ---- fourth batch: BASELINE config C2 on the smallest Transport graph - Omega = ALL lower-triangular edges,
    # X = 0.1 * A_Omega * uniform(0, 1) with seed 4, f = sinh, df = cosh
    An = load_graph("transport_Anaheim")
    An = (An / An.max()).tocsr()
    Tn = sp.tril(An, -1).tocoo()
    order = np.lexsort((Tn.row, Tn.col))
    out["Anaheim"] = sp.csc_matrix(An)
    out["Anaheim_Omega"] = np.stack([Tn.row[order] + 1, Tn.col[order] + 1], 1).astype(np.float64)
    out["Anaheim_X"] = (0.1 * Tn.data[order] * np.random.default_rng(4).uniform(0.0, 1.0, Tn.nnz)).reshape(-1, 1)
    return out


if __name__ == "__main__":
    d = build_inputs()
    path = os.path.join(ROOT, "tests", "golden", "reference_inputs.mat")
    sio.savemat(path, d, format="5", do_compression=True)
    print("wrote", path, os.path.getsize(path), "bytes")
