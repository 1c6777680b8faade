"""The single-launch candidate path (csrc/pairs_small.cuh: one persistent CTA per candidate, whole block-Lanczos
recurrence of functions/trace_fun_update.m:60-125 in one kernel) against the oracle on the shapes its three row zones
and its control flow see: hub rows handled by the whole CTA (>= 1024 nonzeros, Oregon A7), warp rows, 4-lane rows;
a weighted (valued) matrix; iteration caps it = 1, 2, 3 (the reference's lag-2 rule needs j > 2, :104); many more
candidates than CTAs (ticket hand-out).  Tolerance 1e-10 relative, iteration counts and lucky flags equal."""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import edge_UB

pytestmark = pytest.mark.gpu
RTOL = 1e-10


@pytest.fixture(scope="module")
def kr():
    import krylov_robustness_b200 as kr
    return kr


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


def _check(kr, O, A, E, sign, tol, it, fun="exp", every=1):
    n = A.shape[0]
    x, itv, lucky = kr.trace_fun_update_edges(A, E, sign, tol, it, fun)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for h in range(0, len(E), every):
            U, B = edge_UB(n, int(E[h, 0]), int(E[h, 1]), sign)
            ox, oit, olk = O.trace_fun_update(A, U, B, tol, it, 0, fun)
            assert itv[h] == oit and bool(lucky[h]) == bool(olk), (h, itv[h], oit)
            # 1e-10 relative to the value, plus the rounding floor of the formula itself: Xm is a difference of two sums
            # of ~2j terms of size f(||A||) (trace_fun_update.m:85-89), so neither the oracle nor the device can carry
            # more than a few ulps of f(||A||) - it only shows for weak updates (here 3e-5 against sums of size 6)
            floor = 16 * np.finfo(float).eps * tol / 1e-6 if tol > 1e-29 else 0.0
            assert abs(x[h] - ox) <= RTOL * abs(ox) + floor, (h, x[h], ox)
    return x, itv


def test_hub_rows_take_the_whole_cta(kr, O, graphs):
    """Oregon A7: n = 13 947, largest degree 1 761 (a stored row of >= 1024 nonzeros is served by the whole CTA), with
    few enough candidates that the dispatch picks the single-launch path (pairs_small.cuh::use_pairs_small)."""
    A = graphs("oregon_A7")
    assert np.diff(A.indptr).max() >= 1024
    nrm, _ = O.normest(A, 1e-2)
    c = O.compute_centrality(A, "eig")
    E = np.concatenate([O.find_top_edges(A, c, 6, "min"), O.find_top_missing_edges(A, c, 6, "min")])
    _check(kr, O, A, E[:6], -1.0, 1e-6 * float(np.exp(nrm)), 100)
    _check(kr, O, A, E[6:], 1.0, 1e-6 * float(np.exp(nrm)), 100)


@pytest.mark.parametrize("it", [1, 2, 3, 5])
def test_iteration_caps(kr, O, graphs, it):
    A = graphs("oregon_A0")
    c = O.compute_centrality(A, "eig")
    E = O.find_top_edges(A, c, 12, "min")
    x, itv = _check(kr, O, A, E, -1.0, 1e-30, it)           # unreachable tolerance: every candidate runs into the cap
    assert np.all(itv == it)


def test_weighted_graph_and_many_more_candidates_than_ctas(kr, O, graphs):
    """A valued CSR (the Rome road network with random symmetric weights) and 600 candidates on a
    148-SM device: every CTA scores several candidates in turn (ticket counter), results must not depend on it."""
    A0 = graphs("transport_Rome")
    n = A0.shape[0]
    L0 = sp.tril(A0, -1).tocoo()
    w = np.random.default_rng(11).uniform(0.25, 1.0, L0.nnz)
    Lw = sp.coo_matrix((w, (L0.row, L0.col)), shape=(n, n))
    A = (Lw + Lw.T).tocsr()
    assert np.unique(A.data).size > 1
    L = sp.tril(A, -1).tocoo()
    E = np.stack([L.row[:600] + 1, L.col[:600] + 1], 1)
    nrm, _ = O.normest(A, 1e-2)
    tol = 1e-6 * float(np.sinh(nrm))
    x, itv = _check(kr, O, A, E, -0.5, tol, 100, "sinh", every=25)
    again = kr.trace_fun_update_edges(A, E[::-1].copy(), -0.5, tol, 100, "sinh")
    assert np.array_equal(again[0][::-1], x) and np.array_equal(again[1][::-1], itv)
