"""kr_greedy_round with the node-basis screen (csrc/nodepairs.cuh) against the exact round: the selected edge and its
value must be IDENTICAL (the screen only rules candidates out; the selection runs on exact values), the screen's own
values must sit within 1e-6 relative of the exact ones (its documented accuracy is 1e-9 .. 1e-8), and only a small
part of the candidates may take the exact path.  Also the greedy drivers with screen=True reproduce screen=False."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kr():
    import krylov_robustness_b200 as kr
    return kr


@pytest.fixture(scope="module")
def setup(kr):
    import oracle as O
    from krylov_robustness_b200.graphs import power_law_graph, spectral_radius_estimate
    n = 60_000
    A = power_law_graph(n, 1_200_000, 2.2, 7)
    lam = spectral_radius_estimate(A, 30)
    A = (A * (1.0 / lam)).tocsr()
    M = kr.Matrix(A)
    c = kr.compute_centrality(M, "eig")
    E = kr.find_top_missing_edges(A, c, 3000, "min")
    return A, M, lam, E, O


@pytest.mark.parametrize("miobi", ["make", "break"])
def test_screened_round_selects_the_exact_winner(kr, setup, miobi):
    A, M, lam, E, O = setup
    tol = 1e-6 * float(np.e)
    sign = 1.0 if miobi == "make" else -1.0      # 'break' on missing edges is not a real use, but exercises arg-min
    x, it, _ = kr.trace_fun_update_edges(M, E, sign / lam, tol, 100, "exp")
    b0, v0 = kr.select_candidate(x, miobi)
    b1, v1, scores, mask, info = kr.greedy_round(M, E, sign / lam, tol, 100, "exp", miobi, screen=True)
    assert info["screened"] > 0.9 * len(E), info               # the screen settled (almost) everything ...
    assert info["exact"] < 0.2 * len(E), info                  # ... and only the contenders were re-scored
    assert info["nodes"] == np.unique(E).size
    assert (b1, v1) == (b0, v0)
    assert mask[b1]
    assert np.array_equal(scores[mask], x[mask])               # exact where it says so
    rel = np.abs(scores - x) / np.abs(x)
    assert rel.max() <= 1e-6, rel.max()
    b2, v2, s2, m2, info2 = kr.greedy_round(M, E, sign / lam, tol, 100, "exp", miobi, screen=False)
    assert (b2, v2) == (b0, v0) and m2.all() and np.array_equal(s2, x) and info2["screened"] == 0


def test_screened_round_is_checked_against_the_oracle(kr, setup):
    """The winner of the screened round, evaluated by the oracle's trace_fun_update: 1e-10, same step count."""
    import warnings
    from conftest import edge_UB
    A, M, lam, E, O = setup
    tol = 1e-6 * float(np.e)
    b, v, scores, mask, info = kr.greedy_round(M, E[:1200], 1.0 / lam, tol, 100, "exp", "make", screen=True)
    U, B = edge_UB(A.shape[0], int(E[b, 0]), int(E[b, 1]), 1.0 / lam)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ox, oit, _ = O.trace_fun_update(A, U, B, tol, 100)
    assert abs(v - ox) <= 1e-10 * abs(ox)
    # a few screen values against the oracle: inside the screen's documented accuracy
    for h in range(0, 1200, 240):
        U, B = edge_UB(A.shape[0], int(E[h, 0]), int(E[h, 1]), 1.0 / lam)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ox, _, _ = O.trace_fun_update(A, U, B, tol, 100)
        assert abs(scores[h] - ox) <= 1e-7 * abs(ox), (h, scores[h], ox)


def test_greedy_make_with_screen_equals_without(kr, setup):
    A, M, lam, E, O = setup
    c = kr.compute_centrality(M, "eig")
    tol = 1e-6 * float(np.e)
    e0, r0, A0 = kr.greedy_krylov(A, 3, 1500, c, "min", tol, 100, np.inf, 0, "make", rescale=lam)
    e1, r1, A1 = kr.greedy_krylov(A, 3, 1500, c, "min", tol, 100, np.inf, 0, "make", rescale=lam, screen=True)
    assert np.array_equal(e0, e1) and r0 == r1 and (A0 != A1).nnz == 0


def test_sparse_candidate_lists_bypass_the_screen(kr, graphs):
    """'break' candidates on a road network touch as many nodes as there are candidates: not applicable, exact path."""
    import oracle as O
    A = graphs("transport_Rome")
    c = O.compute_centrality(A, "eig")
    E = O.find_top_edges(A, c, 200, "min")
    nrm, _ = O.normest(A, 1e-2)
    b, v, scores, mask, info = kr.greedy_round(A, E, -1.0, 1e-6 * float(np.exp(nrm)), 100, "exp", "break", screen=True)
    assert mask.all() and info["screened"] == 0
    x, _, _ = kr.trace_fun_update_edges(A, E, -1.0, 1e-6 * float(np.exp(nrm)), 100, "exp")
    assert (b, v) == kr.select_candidate(x, "break")


@pytest.mark.parametrize("restrict", ["0", "1"])
def test_screen_gram_modes_agree(kr, setup, monkeypatch, restrict):
    """The cross Grams are computed either for all node pairs (full mode) or only for the column window of the shard's
    second end points (restricted mode, what a rank of a multi-GPU round runs): same winner, same value, screen scores
    equal to the screen's own accuracy."""
    A, M, lam, E, O = setup
    tol = 1e-6 * float(np.e)
    monkeypatch.setenv("KR_SCREEN_RESTRICT", restrict)
    shard = E[2000:3000]                            # a late shard: few second end points, many first ones
    b, v, scores, mask, info = kr.greedy_round(M, shard, 1.0 / lam, tol, 100, "exp", "make", screen=True)
    monkeypatch.delenv("KR_SCREEN_RESTRICT")
    x, _, _ = kr.trace_fun_update_edges(M, shard, 1.0 / lam, tol, 100, "exp")
    assert (b, v) == kr.select_candidate(x, "make")
    assert info["screened"] > 0.9 * len(shard)
    assert np.max(np.abs(scores - x) / np.abs(x)) <= 1e-6
