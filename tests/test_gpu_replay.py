"""End-to-end replays of the reference's experiment scripts as tests (row f4): the SAME optimiser / greedy loop
driven by device callbacks and by oracle callbacks must walk the same path."""
import os
import sys
import warnings

import numpy as np
import pytest

from conftest import ROOT, load_graph

sys.path.insert(0, os.path.join(ROOT, "scripts"))
pytestmark = pytest.mark.gpu


class _Args:
    edges = 8
    search_space = 24
    weight = 10.0
    tol = 1e-6
    it = 100
    maxiter = 60
    hessian = True


@pytest.mark.parametrize("fun,method", [("sinh", "add"), ("exp", "rewire"), ("cosh", "tuning")])
def test_weighted_hessian_experiment_same_path_as_oracle(fun, method):
    """Tests/test_weighted_{sinh,exp,cosh}_hessian.m on a power grid with trust-constr in fmincon's place:
    objective, gradient and Hessian callbacks from the device vs from the oracle - same selected edges, same
    iteration and callback counts, same optimum."""
    import krylov_robustness_b200 as kr
    import oracle as O
    import replay_weighted as R
    A = load_graph("grid_Sweden")
    A = (A / A.max()).tocsr()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d = R.run(kr, A, fun, method, _Args)
        o = R.run(O, A, fun, method, _Args)
    assert d["edges"] == o["edges"]
    assert (d["iterations"], d["callbacks"], d["hessian_callbacks"]) == (o["iterations"], o["callbacks"], o["hessian_callbacks"])
    assert abs(d["fval"] - o["fval"]) <= 1e-9 * abs(o["fval"])
    assert np.max(np.abs(np.array(d["x"]) - np.array(o["x"]))) <= 1e-7


def test_budget_sweep_corner_same_edges_as_oracle():
    """Tests/test_unweighted_break_budget.m call shape (k = 10, Q = 50, road network): identical edge sequence."""
    import krylov_robustness_b200 as kr
    import oracle as O
    A = load_graph("transport_Anaheim")
    nrm = float(np.exp(O.normest(A, 1e-2)[0]))
    c = O.compute_centrality(A, "eig")
    for miobi in ("break", "make"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            e, dtr, _ = kr.greedy_krylov(A, 10, 50, c, "min", 1e-6 * nrm, 100, np.inf, 0, miobi)
            oe, odtr, _ = O.greedy_krylov(A, 10, 50, c, "min", 1e-6 * nrm, 100, np.inf, 0, miobi)
        assert np.array_equal(e, oe)
        assert abs(dtr - odtr) <= 1e-10 * abs(odtr)
