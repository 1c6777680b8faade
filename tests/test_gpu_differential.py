"""Differential test of the DEVICE path against the oracle on the randomised inputs of tests/test_mlab_differential.py
(where the oracle itself is held to the reference's sources): weighted graphs, self loops, rank 1-4 updates, repeated
first indices and diagonal entries, ties in the centrality, iteration caps that are hit, shifted / unshifted expmv with
negative t.  Through the C ABI; 1e-10 relative, counts / flags / edges equal."""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp

from test_mlab_differential import graph, close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kr():
    import krylov_robustness_b200 as kr
    return kr


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


@pytest.mark.parametrize("seed", range(10))
def test_trace_fun_update_random(kr, O, seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(140, 260))
    A = graph(rng, n, weighted=seed % 3 == 1, loops=3 if seed % 4 == 2 else 0)
    rk = int(rng.choice([1, 2, 2, 3, 4]))
    nodes = rng.choice(n, rk, replace=False)
    U = np.zeros((n, rk))
    U[nodes, np.arange(rk)] = 1.0
    B = rng.standard_normal((rk, rk))
    B = (B + B.T) / 2                                          # the device path takes Hermitian B only (it always is)
    fun = ["exp", "sinh", "cosh"][seed % 3]
    it = int(rng.choice([3, 6, 100]))
    tol = 10.0 ** rng.integers(-12, -4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ox, oit, olk = O.trace_fun_update(A, U, B, tol, it, 0, fun)
        x, itr, lk = kr.trace_fun_update(kr.Matrix(A), U, B, tol, it, 0, fun)
    assert itr == oit and bool(lk) == bool(olk), (itr, oit, lk, olk)
    close([x], [ox])


@pytest.mark.parametrize("seed", range(8))
def test_candidate_batches_random_both_pipelines(kr, O, seed, monkeypatch):
    """kr_trace_fun_update_edges on random candidate lists (existing and missing edges, a self loop, a repeated
    candidate) of weighted / unweighted graphs, through the single-launch path (one persistent CTA per candidate) AND
    the batched slot pipeline (KR_PAIR_SLOTS), with iteration caps that some candidates hit."""
    rng = np.random.default_rng(700 + seed)
    n = int(rng.integers(200, 3000))
    A = graph(rng, n, deg=float(rng.choice([3.0, 6.0, 12.0])), weighted=seed % 2 == 1, loops=4 if seed % 4 == 3 else 0)
    m = int(rng.integers(20, 70))
    E = np.stack([rng.integers(1, n + 1, m), rng.integers(1, n + 1, m)], 1).astype(np.int64)
    E[3] = [E[0, 0], E[0, 0]]                                  # self loop (rank-one update, not rescaled)
    E[5] = E[1]                                                # a repeated candidate
    sgn = -1.0 if seed % 2 == 0 else 1.0
    b_off = sgn / (2.0 if seed % 3 == 0 else 1.0)              # rescale = 2 every third case
    it = int(rng.choice([4, 100]))
    nrm = O.normest(A, 1e-2)[0]
    tol = 10.0 ** rng.integers(-9, -5) * float(np.exp(nrm))
    ox, oit, olk = np.zeros(m), np.zeros(m, dtype=np.int64), np.zeros(m, dtype=bool)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for h, (i, j) in enumerate(E):
            if i != j:
                U = np.zeros((n, 2))
                U[i - 1, 0] = U[j - 1, 1] = 1.0
                Bm = b_off * np.array([[0.0, 1.0], [1.0, 0.0]])
            else:
                U = np.zeros((n, 1))
                U[i - 1, 0] = 1.0
                Bm = np.array([[sgn]])
            ox[h], oit[h], olk[h] = O.trace_fun_update(A, U, Bm, tol, it, 0, "exp")
        M = kr.Matrix(A)
        runs = [kr.trace_fun_update_edges(M, E, b_off, tol, it, "exp", b_self=sgn)]
        monkeypatch.setenv("KR_PAIR_SLOTS", "16")
        runs.append(kr.trace_fun_update_edges(M, E, b_off, tol, it, "exp", b_self=sgn))
        monkeypatch.delenv("KR_PAIR_SLOTS")
    deg = np.diff(A.indptr)
    nbr = [set(A.indices[A.indptr[v]:A.indptr[v + 1]].tolist()) | {v} for v in range(n)]
    # rank-deficient first blocks (a leaf end point; adjacent twins) are the documented LAPACK-completion cases
    special = np.array([deg[i - 1] <= 1 or deg[j - 1] <= 1 or (i != j and nbr[i - 1] == nbr[j - 1]) for i, j in E])
    for x, itr, lk in runs:
        assert np.array_equal(itr, oit) and np.array_equal(np.asarray(lk, dtype=bool), olk)
        err = np.abs(x - ox)
        assert np.all(err[~special] <= 1e-10 * np.maximum(np.abs(ox[~special]), 1e-3 * tol)), (err / np.abs(ox)).max()
        assert np.all(err[special] <= 1e-7 * np.maximum(np.abs(ox[special]), tol))


@pytest.mark.parametrize("seed", range(5))
def test_callbacks_and_hessians_random(kr, O, seed):
    rng = np.random.default_rng(200 + seed)
    n = int(rng.integers(300, 420))
    A = graph(rng, n, weighted=True)
    T = sp.tril(A, -1).tocoo()
    pick = np.sort(rng.choice(T.nnz, int(rng.integers(2, 7)), replace=False))
    Om = np.stack([T.row[pick] + 1, T.col[pick] + 1], 1).astype(np.int64)
    X = 0.1 * rng.uniform(0, 1, len(pick))
    M = kr.Matrix(A)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nrm = O.normest(A, 1e-6)[0]
        eA, itE = O.function_multiple_entries(A, Om, "exp", 1e-10 * np.exp(nrm), 100)
        deA, ditE = kr.function_multiple_entries(M, Om, "exp", 1e-10 * np.exp(nrm), 100)
        assert ditE == itE
        close(deA, eA)
        of, ogr = O.fun_and_grad_krylov_exp(X, A, Om, eA, 1e-9, 100, 0)
        f, gr = kr.fun_and_grad_krylov_exp(X, M, Om, eA, 1e-9, 100, 0)
        close(np.concatenate([[f], np.ravel(gr)]), np.concatenate([[of], np.ravel(ogr)]))
        dfA, _ = O.function_multiple_entries(A, Om, "cosh", 1e-10 * np.cosh(nrm), 100)
        of2, ogr2 = O.fun_and_grad_krylov_fun(X, A, Om, "sinh", "cosh", dfA, 1e-9, 100, 0)
        f2, gr2 = kr.fun_and_grad_krylov_fun(X, M, Om, "sinh", "cosh", dfA, 1e-9, 100, 0)
        close(np.ravel(gr2), np.ravel(ogr2))
        assert abs(f2 - of2) <= 1e-8 * abs(of2)               # objective from the Lanczos block: 1e-8 here (rk <= 12, full rank)
        close(kr.hessianfcn_exp(X, A, Om, 1e-9, 100), O.hessianfcn_exp(X, A, Om, 1e-9, 100))
        close(kr.hessianfcn_fun(X, A, Om, "cosh", 1e-9, 100), O.hessianfcn_fun(X, A, Om, "cosh", 1e-9, 100))


@pytest.mark.parametrize("seed", range(6))
def test_entries_with_repeated_rows_and_diagonal(kr, O, seed):
    rng = np.random.default_rng(300 + seed)
    n = int(rng.integers(150, 300))
    A = graph(rng, n, weighted=seed % 2 == 0)
    k = int(rng.integers(3, 12))
    rows = rng.integers(1, n + 1, k)
    rows[k // 2:] = rows[: k - k // 2]
    cols = rng.integers(1, n + 1, k)
    cols[0] = rows[0]
    Om = np.stack([rows, cols], 1).astype(np.int64)
    fun = ["exp", "cosh", "sinh"][seed % 3]
    it = 100 if seed % 3 else 5
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        oX, oit = O.function_multiple_entries(A, Om, fun, 1e-9, it)
        X, itr = kr.function_multiple_entries(kr.Matrix(A), Om, fun, 1e-9, it)
    assert itr == oit
    close(X, oX)


@pytest.mark.parametrize("seed", range(6))
def test_expmv_family_random(kr, O, seed):
    rng = np.random.default_rng(400 + seed)
    n = int(rng.integers(120, 400))
    A = graph(rng, n, deg=6.0, weighted=seed % 2 == 1, loops=5 if seed % 3 == 0 else 0)
    q = int(rng.integers(1, 6))
    b = rng.standard_normal((n, q))
    t = float(rng.choice([1.0, 0.3, 2.5, -1.0]))
    shift = bool(seed % 2 == 0)
    full_term = bool(seed % 3 == 1)
    M = kr.Matrix(A)
    of, os_, om, omv, omvd, ounA = O.expmv(t, A, b, None, "double", shift, False, full_term)
    f, s, m, mv, mvd, unA = kr.expmv(t, M, b, None, "double", shift, False, full_term)
    assert [s, m, mv, mvd, unA] == [os_, om, omv, omvd, ounA]
    close(f, of, 1e-9 if abs(t) > 2 else 1e-10)
    oM, omv2, oal, ounA2 = O.select_taylor_degree(A, b, 55, 8, "double", shift, False, seed % 2 == 1)
    Mt, mv2, al, unA2 = kr.select_taylor_degree(M, b, 55, 8, "double", shift, False, seed % 2 == 1)
    assert mv2 == omv2 and unA2 == ounA2
    close(al, oal, 1e-12)
    close(Mt, oM, 1e-12)
    p = int(rng.integers(2, 8))
    mu = A.diagonal().sum() / n
    for Bm in (A, (A - mu * sp.identity(n)).tocsr()):
        oc, omvn = O.normAm(Bm, p)
        c, mvn = kr.normAm(kr.Matrix(Bm), p)
        assert mvn == omvn
        close([c], [oc], 1e-12)


@pytest.mark.parametrize("seed", range(6))
def test_candidates_and_greedy_random(kr, O, seed):
    rng = np.random.default_rng(600 + seed)
    n = int(rng.integers(140, 220))
    A = graph(rng, n, deg=5.0)
    c = rng.uniform(0.1, 1.0, n)
    if seed % 2:
        c = np.round(c, 1)
    num = int(rng.integers(5, 40))
    for order in ("min", "mult"):
        assert np.array_equal(kr.find_top_edges(A, c, num, order), O.find_top_edges(A, c, num, order))
        assert np.array_equal(kr.find_top_missing_edges(A, c, num, order), O.find_top_missing_edges(A, c, num, order))
    miobi = "break" if seed % 2 == 0 else "make"
    rescale = 1.0 if seed % 3 else 2.0
    tol = 1e-7
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        oe, orob, oA = O.greedy_krylov(A, 3, 12, c, "min", tol, 100, np.inf, 0, miobi, rescale)
        e, rob, An = kr.greedy_krylov(A, 3, 12, c, "min", tol, 100, np.inf, 0, miobi, rescale)
    assert np.array_equal(np.asarray(e), np.asarray(oe)), (e, oe)
    close([rob], [orob])
    assert (sp.csr_matrix(An) != sp.csr_matrix(oA)).nnz == 0
    E = O.find_top_edges(A, c, 6, "mult").astype(np.int64)
    E[2] = [E[0, 0], E[0, 0]]                                  # a self loop among the candidates
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        oe, orob, _ = O.krylov_miobi(A, 2, E, tol, 100, np.inf, 0, miobi, rescale)
        e, rob, _ = kr.krylov_miobi(A, 2, E, tol, 100, np.inf, 0, miobi, rescale)
    assert np.array_equal(np.asarray(e), np.asarray(oe)), (e, oe)
    close([rob], [orob])
