"""GPU parity: SpMM and the throughput-mode SLQ estimator against the oracle (through the C ABI)."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kr():
    import krylov_robustness_b200 as kr
    return kr


@pytest.mark.parametrize("gname", ["oregon_A0", "oregon_A8", "transport_Rome", "grid_England"])
@pytest.mark.parametrize("k", [1, 2, 8, 10, 37])
def test_spmm_matches_scipy(kr, graphs, gname, k):
    A = graphs(gname)
    n = A.shape[0]
    X = np.random.default_rng(k).standard_normal((n, k))
    M = kr.Matrix(A)
    Y = M @ X
    ref = A @ X
    scale = (abs(A) @ np.abs(X)).max()
    assert np.abs(Y - ref).max() <= 1e-14 * scale           # fp64, differs only by summation order
    info = M.info()
    assert info["symmetric"] and info["nnz"] == A.nnz
    assert info["pattern_only"] == gname.startswith(("oregon", "transport"))


def test_spmm_unsymmetric_ragged_and_empty_rows(kr):
    rng = np.random.default_rng(0)
    n = 3000
    A = sp.random(n, n, density=0.002, random_state=1, format="lil")
    A[5, :] = 0                       # empty row
    A[7, :] = rng.standard_normal(n)  # one full row (n nonzeros)
    A[:, 11] = 0                      # empty column
    A = A.tocsr()
    X = rng.standard_normal((n, 5))
    M = kr.Matrix(A)
    assert not M.info()["symmetric"]
    Y = M @ X
    ref = A @ X
    assert np.abs(Y - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.all(Y[5] == 0)


def test_spmm_errors(kr, graphs):
    M = kr.Matrix(graphs("oregon_A0"))
    with pytest.raises(ValueError, match="wrong number of rows"):
        M @ np.ones((5, 2))
    with pytest.raises(ValueError, match="should be square"):
        kr.Matrix(sp.csr_matrix(np.ones((3, 4))))


def test_dense_roundtrip_and_rademacher(kr):
    X = np.random.default_rng(3).standard_normal((1000, 13))
    D = kr.Dense(1000, 13).upload(X)
    assert np.array_equal(D.download(), X)
    R = kr.Dense(777, 9).fill_rademacher(42, col_offset=5).download()
    assert np.array_equal(R, kr.rademacher_host(777, 9, 42, col_offset=5))
    assert set(np.unique(R)) == {-1.0, 1.0}


@pytest.mark.parametrize("gname,fun", [("oregon_A0", "exp"), ("transport_Rome", "exp"), ("grid_Mexico", "cosh"),
                                       ("oregon_A8", "sinh")])
def test_slq_matches_oracle(kr, graphs, gname, fun):
    import oracle as O
    A = graphs(gname)
    if gname.startswith("oregon"):
        A = A / 8.0                                         # keep exp in a comfortable range
    n = A.shape[0]
    Z = kr.rademacher_host(n, 24, 7)
    tr, vals, al, be = kr.slq_trace(A, Z, 20, fun, return_details=True)
    otr, ovals, oal, obe = O.slq_trace(A, Z, 20, fun)
    # fp64 tolerance of the north star: 1e-10 relative on traces
    assert abs(tr - otr) <= 1e-10 * abs(otr)
    assert np.max(np.abs(vals - ovals)) <= 1e-10 * np.max(np.abs(ovals))
    assert np.max(np.abs(al - oal)) <= 1e-9 * np.max(np.abs(oal))
    # device-resident probes give the identical result (same kernels, same order)
    D = kr.Dense(n, 24).upload(Z)
    assert kr.slq_trace(kr.Matrix(A), D, 20, fun) == tr
    assert kr.slq_trace(kr.Matrix(A), D, 20, fun) == tr      # bit-reproducible run to run
    # int8 sign probes (kr_slq_trace_sign): 1/8 of the upload, same fp64 arithmetic, identical result;
    # a ragged leading dimension exercises the strided copy
    Zs = np.zeros((n + 5, 24), dtype=np.int8, order="F")
    Zs[:n] = Z.astype(np.int8)
    tr8, vals8, _, _ = kr.slq_trace(A, Zs[:n], 20, fun, return_details=True)
    assert tr8 == tr and np.array_equal(vals8, vals)


def test_slq_power_law_64_probes(kr):
    """A mid-size instance of the bench workload (Chung-Lu power law, hubs + long rows > 1024 nonzeros
    exercise the CTA-per-row path) against the oracle."""
    import oracle as O
    from krylov_robustness_b200.graphs import power_law_graph, spectral_radius_estimate
    A = power_law_graph(60_000, 2_400_000, 2.1, seed=5)
    assert np.diff(A.indptr).max() >= 1024
    A = (A * (1.0 / spectral_radius_estimate(A, 30))).tocsr()
    n = A.shape[0]
    Z = kr.rademacher_host(n, 40, 11)
    tr, vals, al, be = kr.slq_trace(A, Z, 25, "exp", return_details=True)
    otr, ovals, oal, obe = O.slq_trace(A, Z, 25, "exp")
    assert abs(tr - otr) <= 1e-10 * abs(otr)
    assert np.max(np.abs(vals - ovals)) <= 1e-10 * np.max(np.abs(ovals))
    X = np.random.default_rng(1).standard_normal((n, 21))
    assert np.abs(kr.Matrix(A) @ X - A @ X).max() <= 1e-13 * np.abs(X).max() * 50


def test_empty_and_degenerate_inputs(kr, graphs):
    A = graphs("oregon_A0")
    M = kr.Matrix(A)
    x, it, lucky = kr.trace_fun_update_edges(M, np.zeros((0, 2), dtype=np.int64), -1.0, 1e-3, 100, "exp")
    assert x.size == 0 and it.size == 0
    X, it = kr.function_multiple_entries(M, np.zeros((0, 2), dtype=np.int64), "exp", 1e-8, 50)
    assert X.size == 0 and it == 0
    with pytest.raises(kr._lib.KrylovB200Error, match="out of range"):
        kr.trace_fun_update_edges(M, np.array([[1, 700]]), -1.0, 1e-3, 100, "exp")
    with pytest.raises(ValueError, match="unsupported function handle"):
        kr.trace_fun_update_edges(M, np.array([[1, 2]]), -1.0, 1e-3, 100, np.tanh)
    import scipy.sparse as sp
    Mz = kr.Matrix(sp.csr_matrix((300, 300)))          # all-zero matrix: empty rows everywhere
    Y = Mz @ np.ones((300, 3))
    assert np.all(Y == 0)
    assert Mz.info()["nnz"] == 0


def test_slq_breakdown_and_duplicates(kr):
    """Lanczos breakdown (beta = 0: probe supported on an isolated node / an eigenvector) and a CSR input
    with duplicate, unsorted entries (summed like MATLAB's sparse())."""
    import oracle as O
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    n = 400
    B = sp.random(n - 3, n - 3, density=0.02, random_state=2)
    B = ((B + B.T) * 0.5).tocsr()
    A = sp.block_diag([B, sp.csr_matrix((3, 3))], format="csr")     # three isolated nodes
    Z = rng.standard_normal((n, 5))
    Z[:, 0] = 0.0
    Z[n - 1, 0] = 2.0                                               # A z = 0: breakdown at step 1
    tr, vals, al, be = kr.slq_trace(A, Z, 12, "exp", return_details=True)
    otr, ovals, oal, obe = O.slq_trace(A, Z, 12, "exp")
    assert vals[0] == 4.0 == ovals[0] and be[0, 0] == 0.0
    assert np.max(np.abs(vals - ovals)) <= 1e-10 * np.max(np.abs(ovals))
    # duplicates + unsorted columns in the raw CSR arrays
    rows = np.array([0, 0, 0, 1, 2, 2]); cols = np.array([2, 1, 2, 0, 0, 0]); data = np.array([1.0, 3.0, 0.5, 3.0, 1.0, 0.5])
    indptr = np.array([0, 3, 4, 6]); 
    Araw = sp.csr_matrix((data, cols, indptr), shape=(3, 3))        # not canonical on purpose
    M = kr.Matrix.__new__(kr.Matrix)
    import ctypes as C
    ctx = kr.Context.default()
    h = C.c_void_p()
    rp = indptr.astype(np.int64); ci = cols.astype(np.int64)
    kr._lib.check(ctx.lib.kr_matrix_create(ctx.h, 3, 6, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p),
                                           data.ctypes.data_as(C.c_void_p), C.byref(h)))
    M.ctx, M.h, M.n, M.shape = ctx, h, 3, (3, 3)
    X = np.eye(3)
    dense = np.zeros((3, 3))
    for r in range(3):
        for p in range(indptr[r], indptr[r + 1]):
            dense[r, cols[p]] += data[p]
    assert np.array_equal(M @ X, dense)
    assert M.info()["symmetric"] and M.info()["nnz"] == 4
