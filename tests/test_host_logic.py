"""Host-side logic of the candidate generators (no GPU needed): krylov_robustness_b200.functions against the oracle's
restatement of functions/find_top_edges.m / find_top_missing_edges.m on the reference's graphs."""
import numpy as np
import pytest

from conftest import load_graph


@pytest.mark.parametrize("gname", ["oregon_A0", "transport_Anaheim", "grid_Mexico"])
@pytest.mark.parametrize("order", ["min", "mult"])
def test_find_top_missing_edges_matches_oracle(gname, order):
    import oracle as O
    from krylov_robustness_b200 import functions as F
    A = load_graph(gname)
    A = (A != 0).astype(np.float64).tocsr()
    rng = np.random.default_rng(7)
    for c in (np.asarray(A.sum(axis=0)).ravel() + 1e-3 * rng.random(A.shape[0]),       # distinct values
              np.asarray(A.sum(axis=0)).ravel()):                                      # ties (degree centrality)
        for num in (1, 17, 250):
            E = F.find_top_missing_edges(A, c, num, order)
            Eo = O.find_top_missing_edges(A, c, num, order)
            assert np.array_equal(E, Eo), (gname, order, num)
            assert all(A[i - 1, j - 1] == 0 and i != j for i, j in E)


def test_find_top_missing_edges_rejects_unknown_order():
    from krylov_robustness_b200 import functions as F
    A = load_graph("oregon_A0")
    with pytest.raises(ValueError):
        F.find_top_missing_edges(A, np.ones(A.shape[0]), 3, "max")


def test_compute_centrality_degree_and_res():
    from krylov_robustness_b200 import functions as F
    A = load_graph("oregon_A0")
    assert np.array_equal(F.compute_centrality(A, "deg"), np.asarray(A.sum(axis=0)).ravel())
    with pytest.raises(NameError):
        F.compute_centrality(A, "res")      # functions/compute_centrality.m:14 uses an undefined n
