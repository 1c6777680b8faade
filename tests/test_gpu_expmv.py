"""GPU parity of the expmv family and mc_trace against the oracle (through the C ABI)."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kr():
    import krylov_robustness_b200 as kr
    return kr


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


def test_theta_and_normAm(kr, O, graphs):
    import ctypes as C
    th = (C.c_double * 100)()
    kr._lib.check(kr._lib.load().kr_theta(th))
    assert np.array_equal(np.array(th), O.THETA)
    A = graphs("oregon_A0")
    for m in (1, 2, 5, 9):
        c, mv = kr.normAm(A, m)
        oc, omv = O.normAm(A, m)
        assert mv == omv == m and abs(c - oc) <= 1e-13 * oc
    # signed matrix: the 1-norm estimator branch (normAm.m:25-26)
    B = graphs("grid_England").copy()
    B.data[::3] *= -1
    B = ((B + B.T) / 2).tocsr()
    c, mv = kr.normAm(B, 3)
    oc, omv = O.normAm(B, 3)
    assert mv == omv and abs(c - oc) <= 1e-12 * oc


@pytest.mark.parametrize("gname,q", [("oregon_A0", 10), ("oregon_A8", 3), ("transport_Rome", 16), ("grid_Mexico", 1)])
def test_expmv_vs_oracle(kr, O, graphs, gname, q):
    A = graphs(gname)
    n = A.shape[0]
    b = np.sign(np.random.default_rng(q).standard_normal((n, q)))
    M, mv, alpha, unA = kr.select_taylor_degree(A, b)
    oM, omv, oalpha, ounA = O.select_taylor_degree(A, b)
    assert (mv, unA) == (omv, ounA)
    assert np.allclose(M, oM, rtol=1e-13, atol=0) and np.allclose(alpha, oalpha, rtol=1e-13)
    f, s, m, mv, mvd, unA = kr.expmv(1, A, b)
    of, os_, om, omv, omvd, ounA = O.expmv(1, A, b)
    assert (s, m, mv, mvd, unA) == (os_, om, omv, omvd, ounA)
    assert np.linalg.norm(f - of) <= 1e-12 * np.linalg.norm(of)
    # precomputed M, t != 1, full_term
    f2, s2, m2, mv2, mvd2, _ = kr.expmv(0.5, A, b, M=oM, full_term=True)
    of2, os2, om2, omv2, omvd2, _ = O.expmv(0.5, A, b, M=oM, full_term=True)
    assert (s2, m2, mv2, mvd2) == (os2, om2, omv2, omvd2)
    assert np.linalg.norm(f2 - of2) <= 1e-12 * np.linalg.norm(of2)
    assert np.array_equal(kr.expmv(0, A, b)[0], b)


def test_expmv_shift_with_diagonal(kr, O, graphs):
    A = (graphs("grid_Mexico") + sp.diags(np.linspace(0.5, 1.5, 552))).tocsr()
    b = np.random.default_rng(2).standard_normal((552, 4))
    f, s, m, mv, mvd, unA = kr.expmv(0.7, A, b)
    of, os_, om, omv, omvd, ounA = O.expmv(0.7, A, b)
    assert (s, m, mv, mvd, unA) == (os_, om, omv, omvd, ounA)
    assert np.linalg.norm(f - of) <= 1e-12 * np.linalg.norm(of)
    with pytest.raises(ValueError, match="Invalid p_max or m_max"):
        kr.select_taylor_degree(A, b, 61, 8)


def test_mc_trace_and_trace_exp_vs_oracle(kr, O, graphs):
    A = graphs("oregon_A0")
    n = A.shape[0]
    rng = np.random.default_rng(0)
    probes = [(np.sign(rng.standard_normal((n, 10))), np.sign(rng.standard_normal((n, 10)))) for _ in range(34)]
    tr, res, it = kr.mc_trace(kr.ExpmvHandle(A), n, 1e-4, 1000, 1, probes=probes)
    otr, ores, oit = O.mc_trace(lambda x: O.expmv(1, A, x)[0], n, 1e-4, 1000, 1, probes=probes)
    assert it == oit
    assert abs(tr - otr) <= 1e-9 * abs(otr)
    assert abs(kr.trace_exp(A, probes) - O.trace_exp(A, probes=probes)) <= 1e-9 * abs(otr)
    # numeric Afun (mc_trace.m:32-34)
    As = (A / 20.0).tocsr()
    tr2, res2, it2 = kr.mc_trace(As, n, 1e-3, 90, 1, probes=probes)
    otr2, ores2, oit2 = O.mc_trace(As, n, 1e-3, 90, 1, probes=probes)
    assert it2 == oit2 and abs(tr2 - otr2) <= 1e-9 * max(1.0, abs(otr2))


def test_full_size_spmm_properties(kr):
    """BASELINE config C3 at full size (n = 1M, nnz = 20M): direct parity on 16 columns plus the
    size-independent properties (linearity, symmetry x'(Ay) = y'(Ax))."""
    from krylov_robustness_b200.graphs import power_law_graph
    A = power_law_graph(1_000_000, 20_000_000, 2.2, 20260310)
    n = A.shape[0]
    M = kr.Matrix(A)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, 16))
    Y = M @ X
    ref = A @ X
    assert np.abs(Y - ref).max() <= 1e-13 * np.abs(ref).max()
    Z = rng.standard_normal((n, 16))
    assert np.abs(M @ (2.0 * X + Z) - (2.0 * Y + (M @ Z))).max() <= 1e-12 * np.abs(Y).max()
    assert abs(np.vdot(X[:, 0], (M @ Z)[:, 0]) - np.vdot(Z[:, 0], Y[:, 0])) <= 1e-9 * np.abs(Y).max() * np.sqrt(n)


def test_expmv_rmat_64_columns(kr, O):
    """Config C4 shape at a size the oracle finishes in seconds: R-MAT graph, 64 right-hand sides,
    normAm-driven degree selection on the device, scaled so that exp stays in range."""
    from krylov_robustness_b200.graphs import rmat_graph, spectral_radius_estimate
    A = rmat_graph(scale=15, nnz=1 << 19, seed=2)
    lam = spectral_radius_estimate(A, 30)
    A = (A * (4.0 / lam)).tocsr()
    n = A.shape[0]
    b = np.random.default_rng(3).standard_normal((n, 64))
    f, s, m, mv, mvd, unA = kr.expmv(1, A, b)
    of, os_, om, omv, omvd, ounA = O.expmv(1, A, b)
    assert (s, m, mv, mvd, unA) == (os_, om, omv, omvd, ounA)
    assert np.linalg.norm(f - of) <= 1e-12 * np.linalg.norm(of)
    # linearity in b (size-independent property)
    f2 = kr.expmv(1, A, 3.0 * b[:, :5])[0]
    assert np.linalg.norm(f2 - 3.0 * f[:, :5]) <= 1e-12 * np.linalg.norm(f2)
