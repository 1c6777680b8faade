"""Committed golden vectors (tests/golden/oracle_golden.json, scripts/make_golden.py).

CPU part: the oracle reproduces its own pinned outputs (guards against drift of the checker).
GPU part: the device path hits the same numbers through the C ABI without re-running the oracle.
The reference itself ships no golden vectors; outputs of the reference's own sources are pinned separately in
tests/test_reference_goldens.py (tests/golden/reference_golden.json)."""
import json
import os
import warnings

import numpy as np
import pytest

from conftest import GOLDEN, edge_UB, load_graph

G = json.load(open(os.path.join(GOLDEN, "oracle_golden.json")))
RTOL = 1e-10


def _edge_cases():
    return [k for k in G if k.startswith("edges/")]


@pytest.mark.parametrize("key", _edge_cases())
def test_oracle_reproduces_edge_goldens(key):
    import oracle as O
    _, gname, fun, sign = key.split("/")
    sign = float(sign)
    A = load_graph(gname)
    g = G[key]
    for (i, j), x, it in list(zip(g["E"], g["Xm"], g["iter"]))[:6]:
        U, B = edge_UB(A.shape[0], i, j, sign)
        ox, oit, _ = O.trace_fun_update(A, U, B, g["tol"], 100, 0, fun)
        assert oit == it and abs(ox - x) <= 1e-12 * abs(x)


def test_oracle_reproduces_expmv_and_entries_goldens():
    import oracle as O
    for key in [k for k in G if k.startswith("expmv/")]:
        _, gname, q = key.split("/")
        q = int(q)
        A = load_graph(gname)
        b = np.sign(np.random.default_rng(q).standard_normal((A.shape[0], q)))
        f, s, m, mv, mvd, unA = O.expmv(1, A, b)
        g = G[key]
        assert (s, m, mv, mvd, unA) == (g["s"], g["m"], g["mv"], g["mvd"], g["unA"])
        assert abs(np.linalg.norm(f) - g["fro"]) <= 1e-12 * g["fro"]
    g = G["entries/oregon_A0/exp"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        X, it = O.function_multiple_entries(load_graph("oregon_A0"), np.array(g["omega"]), "exp", g["tol"], 100)
    assert it == g["iter"] and np.max(np.abs(X - np.array(g["X"]))) <= 1e-12 * np.max(np.abs(g["X"]))


# ------------------------------------------------------------------ device vs goldens
@pytest.fixture(scope="module")
def kr():
    import krylov_robustness_b200 as kr
    return kr


@pytest.mark.gpu
@pytest.mark.parametrize("key", _edge_cases())
def test_device_hits_edge_goldens(kr, key):
    _, gname, fun, sign = key.split("/")
    g = G[key]
    x, it, _ = kr.trace_fun_update_edges(load_graph(gname), np.array(g["E"]), float(sign), g["tol"], 100, fun)
    assert list(it) == g["iter"]
    assert np.max(np.abs(x - np.array(g["Xm"])) / np.abs(g["Xm"])) <= RTOL


@pytest.mark.gpu
def test_device_hits_expmv_entries_slq_goldens(kr):
    for key in [k for k in G if k.startswith("expmv/")]:
        _, gname, q = key.split("/")
        q = int(q)
        A = load_graph(gname)
        b = np.sign(np.random.default_rng(q).standard_normal((A.shape[0], q)))
        f, s, m, mv, mvd, unA = kr.expmv(1, A, b)
        g = G[key]
        assert (s, m, mv, mvd, unA) == (g["s"], g["m"], g["mv"], g["mvd"], g["unA"])
        assert np.max(np.abs(f.sum(axis=0) - np.array(g["colsum"]))) <= 1e-11 * g["fro"]
    g = G["entries/oregon_A0/exp"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        X, it = kr.function_multiple_entries(load_graph("oregon_A0"), np.array(g["omega"]), "exp", g["tol"], 100)
    assert it == g["iter"] and np.max(np.abs(X - np.array(g["X"]))) <= RTOL * np.max(np.abs(g["X"]))
    g = G["slq/oregon_A8_div8/exp"]
    A = (load_graph("oregon_A8") / 8.0).tocsr()
    tr, vals, _, _ = kr.slq_trace(A, kr.rademacher_host(A.shape[0], g["k"], g["seed"]), g["m"], "exp", return_details=True)
    assert abs(tr - g["tr"]) <= RTOL * abs(g["tr"])
    assert np.max(np.abs(vals - np.array(g["vals"]))) <= RTOL * np.max(np.abs(g["vals"]))


@pytest.mark.gpu
def test_device_hits_gradient_hessian_greedy_goldens(kr):
    g = G["grad_exp/oregon_A1"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f, gr = kr.fun_and_grad_krylov_exp(np.array(g["X"]), load_graph("oregon_A1"), np.array(g["Omega"]),
                                           np.array(g["eA"]), 1e-8, 100)
        assert abs(f - g["f"]) <= RTOL * abs(g["f"])
        assert np.linalg.norm(gr - np.array(g["gr"])) <= RTOL * np.linalg.norm(g["gr"])
        g = G["hessian_exp/grid_England"]
        H = kr.hessianfcn_exp(np.array(g["X"]), load_graph("grid_England"), np.array(g["Omega"]), 1e-10, 60)
        assert np.max(np.abs(H - np.array(g["Hes"]))) <= RTOL * np.max(np.abs(g["Hes"]))
        A = load_graph("transport_Barcelona")
        for miobi in ("break", "make"):
            g = G["greedy/transport_Barcelona/%s" % miobi]
            e, rob, _ = kr.greedy_krylov(A, 4, 30, np.array(g["centrality"]), "min", g["tol"], 100, np.inf, 0, miobi)
            assert e.tolist() == g["edges"] and abs(rob - g["rob"]) <= RTOL * abs(g["rob"])
