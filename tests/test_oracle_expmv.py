"""Oracle: expmv / select_taylor_degree / normAm / mc_trace / slq_trace / greedy drivers."""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp

import oracle as O


def test_theta_table():
    # values quoted in SURVEY.md App. C / Al-Mohy & Higham 2011 Table 3.1
    assert O.THETA[0] == 2.2204460492503131e-16
    assert O.THETA[54] == 9.8674966757534008
    assert np.all(np.diff(O.THETA) > 0)


def test_normAm_exact_for_nonnegative(graphs):
    A = graphs("oregon_A0")
    Ad = A.toarray()
    for m in (1, 2, 5, 9):
        c, mv = O.normAm(A, m)
        assert mv == m
        assert np.isclose(c, np.linalg.norm(np.linalg.matrix_power(Ad, m), 1), rtol=1e-13)


def test_normAm_general_branch_is_lower_bound(graphs):
    A = graphs("grid_England").copy()
    A.data[::3] *= -1
    A = ((A + A.T) / 2).tocsr()
    c, mv = O.normAm(A, 3)
    exact = np.linalg.norm(np.linalg.matrix_power(A.toarray(), 3), 1)
    assert mv % 3 == 0 and mv >= 6
    assert 0.3 * exact <= c <= exact * (1 + 1e-12)


def test_select_taylor_degree_and_expmv_counts(graphs):
    A = graphs("oregon_A0")
    S = np.sign(np.random.default_rng(0).standard_normal((A.shape[0], 10)))
    M, mv, alpha, unA = O.select_taylor_degree(A, S)
    assert M.shape == (55, 7) and mv == 44 and unA == 0
    f, s, m, mv, mvd, unA = O.expmv(1, A, S)
    # sizing probe recorded in SURVEY.md 3.1 / BASELINE.md 2
    assert (m, s, mv, mvd) == (52, 3, 164, 44)
    ref = sla.expm(A.toarray()) @ S
    assert np.linalg.norm(f - ref) <= 1e-12 * np.linalg.norm(ref)
    with pytest.raises(ValueError, match="Invalid p_max or m_max"):
        O.select_taylor_degree(A, S, 61, 8)


def test_expmv_small_norm_uses_normA(graphs):
    A = graphs("grid_Sweden") * 0.05
    b = np.random.default_rng(1).standard_normal((A.shape[0], 3))
    f, s, m, mv, mvd, unA = O.expmv(1, A, b)
    assert unA == 1 and mvd == 0 and s == 1
    ref = sla.expm(A.toarray()) @ b
    assert np.linalg.norm(f - ref) <= 1e-13 * np.linalg.norm(ref)


def test_expmv_shift_weighted_diagonal(graphs):
    A = graphs("grid_Mexico") + sp.diags(np.linspace(0.5, 1.5, 552))
    b = np.random.default_rng(2).standard_normal((552, 4))
    f = O.expmv(0.7, A.tocsr(), b)[0]
    ref = sla.expm(0.7 * A.toarray()) @ b
    assert np.linalg.norm(f - ref) <= 1e-12 * np.linalg.norm(ref)
    assert np.array_equal(O.expmv(0, A.tocsr(), b)[0], b)


def test_mc_trace_matches_dense_trace(graphs):
    A = graphs("oregon_A0")
    n = A.shape[0]
    rng = np.random.default_rng(0)
    probes = [(np.sign(rng.standard_normal((n, 10))), np.sign(rng.standard_normal((n, 10)))) for _ in range(34)]
    tr, res, it = O.mc_trace(lambda x: O.expmv(1, A, x)[0], n, 1e-4, 1000, 1, probes=probes)
    true = np.exp(np.linalg.eigvalsh(A.toarray())).sum()
    assert it >= 2 and res < 1e-4
    assert abs(tr - true) <= 1e-4 * true
    # numeric Afun form (mc_trace.m:32-34)
    Ad = A.toarray() / 20
    tr2, _, _ = O.mc_trace(Ad, n, 1e-3, 300, 1, probes=probes)
    assert abs(tr2 - np.trace(Ad)) < 1.0


def test_slq_trace_against_dense(graphs):
    A = graphs("transport_Rome")
    n = A.shape[0]
    Z = np.sign(np.random.default_rng(1).standard_normal((n, 256)))
    tr, vals, alpha, beta = O.slq_trace(A, Z, 30, "exp")
    true = np.exp(np.linalg.eigvalsh(A.toarray())).sum()
    assert abs(tr - true) <= 0.05 * true            # Monte-Carlo error, 256 probes
    # per-probe value is the Gauss quadrature of z' f(A) z
    z = Z[:, 0]
    assert abs(vals[0] - z @ sla.expm(A.toarray()) @ z) <= 1e-8 * abs(vals[0])


def test_find_top_edges_orders(graphs):
    A = graphs("oregon_A0")
    c = O.compute_centrality(A, "eig")
    E = O.find_top_edges(A, c, 50, "mult")
    prod = c[E[:, 0] - 1] * c[E[:, 1] - 1]
    assert np.all(np.diff(prod) <= 1e-18) and np.all(E[:, 0] > E[:, 1])
    assert all(A[i - 1, j - 1] == 1 for i, j in E)
    Em = O.find_top_edges(A, c, 50, "min")
    assert len({tuple(e) for e in Em}) == 50
    Mi = O.find_top_missing_edges(A, c, 40, "min")
    assert all(A[i - 1, j - 1] == 0 and i != j for i, j in Mi)
    rank = {node: r for r, node in enumerate(np.argsort(-c, kind="stable"))}
    worst = [max(rank[i - 1], rank[j - 1]) for i, j in Mi]
    assert worst == sorted(worst)


def test_greedy_break_and_make_vs_dense(graphs):
    A = graphs("transport_Anaheim")
    nrm, _ = O.normest(A, 1e-2)
    c = O.compute_centrality(A, "eig")
    tol = 1e-6 * np.exp(nrm)
    tr0 = np.exp(np.linalg.eigvalsh(A.toarray())).sum()
    e, rob, An = O.greedy_krylov(A, 4, 30, c, "min", tol, 100, np.inf, 0, "break")
    assert e.shape == (4, 2) and An.nnz == A.nnz - 8
    assert abs((np.exp(np.linalg.eigvalsh(An.toarray())).sum() - tr0) - rob) <= 1e-5 * abs(rob)
    e, rob, An = O.greedy_krylov(A, 3, 30, c, "min", tol, 100, np.inf, 0, "make")
    assert An.nnz == A.nnz + 6
    assert abs((np.exp(np.linalg.eigvalsh(An.toarray())).sum() - tr0) - rob) <= 1e-5 * abs(rob)
    with pytest.raises(ValueError, match="should be symmetric"):
        O.greedy_krylov(sp.triu(A).tocsr(), 1, 5, c)
    with pytest.raises(ValueError, match="more than edges"):
        O.greedy_krylov(A, A.nnz, 5, c)


def test_theta_and_expmv_against_scipys_independent_implementation(graphs):
    """An independent anchor for the expmv family: SciPy's expm_multiply implements the same Al-Mohy - Higham 2011
    algorithm from the paper, with its own copy of the theta table.  (1) The theta values the reference ships in
    functions/theta_taylor.mat (oracle/theta.py, csrc/theta_table.h) equal SciPy's for every degree SciPy
    tabulates; (2) oracle.expmv and scipy agree on e^{tA} B to 1e-12 on the reference's graphs."""
    from scipy.sparse.linalg import expm_multiply
    from scipy.sparse.linalg import _expm_multiply as em
    from oracle.theta import THETA
    import oracle as O
    for m, th in em._theta.items():
        # SciPy prints 2-3 significant digits; theta_1 is 2^-52 in the .mat file and 2.29e-16 in the paper's table
        assert abs(THETA[m - 1] - th) <= 5e-2 * th, (m, THETA[m - 1], th)
    for gname, t in (("oregon_A0", 0.25), ("transport_Rome", 1.0), ("grid_Mexico", 1.0)):
        A = graphs(gname)
        b = np.random.default_rng(5).standard_normal((A.shape[0], 3))
        f = O.expmv(t, A, b)[0]
        g = expm_multiply(t * A, b)
        assert np.max(np.abs(f - g)) <= 1e-12 * np.max(np.abs(g))
