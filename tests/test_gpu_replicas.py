"""Multi-GPU inside one process (kr_matrix_replicate): A replicated on every visible GPU, candidate
edges / probe columns split across the replicas by host threads behind the C ABI.  On a one-GPU box the replica
list is empty and the calls degenerate to the single-GPU path (still exercised); with >= 2 GPUs the sharded results
must equal the single-GPU ones: per-candidate values bit for bit (a candidate's arithmetic does not depend on which
GPU or slot runs it), the trace to rounding of the final sum."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kr():
    import krylov_robustness_b200 as kr
    return kr


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_replicated_candidate_scoring_and_greedy_edits(kr, graphs):
    import oracle as O
    A = graphs("oregon_A8")
    nrm, _ = O.normest(A, 1e-2)
    tol = 1e-6 * float(np.exp(nrm))
    c = O.compute_centrality(A, "eig")
    E = O.find_top_edges(A, c, 96, "min")
    single = kr.Matrix(A)
    x0, it0, lk0 = kr.trace_fun_update_edges(single, E, -1.0, tol, 100, "exp")
    multi = kr.Matrix(A)
    R = multi.replicate()
    assert R == max(_ngpu(), 1) == multi.replicas()
    x1, it1, lk1 = kr.trace_fun_update_edges(multi, E, -1.0, tol, 100, "exp")
    assert np.array_equal(it0, it1) and np.array_equal(lk0, lk1)
    assert np.array_equal(x0, x1)
    # an edge edit reaches every replica: delete the best edge, rescore, compare with the single-GPU matrix
    best = int(np.argmin(x0))
    for M in (single, multi):
        M.set_edges([E[best, 0]], [E[best, 1]], 0.0)
    Er = np.delete(E, best, axis=0)
    y0 = kr.trace_fun_update_edges(single, Er, -1.0, tol, 100, "exp")
    y1 = kr.trace_fun_update_edges(multi, Er, -1.0, tol, 100, "exp")
    assert np.array_equal(y0[0], y1[0]) and np.array_equal(y0[1], y1[1])
    if R > 1:
        with pytest.raises(kr._lib.KrylovB200Error, match="already has replicas"):
            multi.replicate()


def test_replicated_slq_trace_splits_probe_columns(kr, graphs):
    A = graphs("oregon_A0") / 8.0
    n = A.shape[0]
    Z = kr.rademacher_host(n, 64, 5)
    single = kr.Matrix(A)
    t0, v0, a0, b0 = kr.slq_trace(single, Z, 12, "exp", return_details=True)
    multi = kr.Matrix(A)
    multi.replicate()
    t1, v1, a1, b1 = kr.slq_trace(multi, Z, 12, "exp", return_details=True)
    assert abs(t1 - t0) <= 1e-13 * abs(t0)
    # probe columns are independent: the per-column quadratures agree to rounding (the Lanczos dots of a column do
    # not depend on its position in a panel)
    assert np.allclose(v1, v0, rtol=1e-12, atol=0) and np.allclose(a1, a0, rtol=1e-12, atol=1e-14)
    zi = Z.astype(np.int8)
    assert abs(kr.slq_trace(multi, zi, 12, "exp") - t0) <= 1e-13 * abs(t0)


def test_kr_gpus_environment_replicates_at_creation(kr, graphs, monkeypatch):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    monkeypatch.setenv("KR_GPUS", "2")
    M = kr.Matrix(graphs("oregon_A0"))
    assert M.replicas() == 2
