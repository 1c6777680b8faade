"""The route to PINNED parity.

The reference is MATLAB-only and ships no golden vectors; this image has neither MATLAB nor Octave, so nothing here can
run the reference.  scripts/make_reference_goldens.m runs the reference's OWN functions/*.m on the committed inputs
tests/golden/reference_inputs.mat (made by scripts/make_reference_inputs.py) and writes
tests/golden/reference_golden.json.  When that file is present these tests check the oracle (CPU tier) and the device
path (GPU tier) against it at 1e-10 with equal iteration counts; until then they are skipped and parity stays
"unpinned"."""
import json
import os
import warnings

import numpy as np
import pytest
import scipy.io as sio
import scipy.sparse as sp

from conftest import GOLDEN, ROOT, edge_UB

REF = os.path.join(GOLDEN, "reference_golden.json")
RTOL = 1e-10


def inputs():
    d = sio.loadmat(os.path.join(GOLDEN, "reference_inputs.mat"))
    return {k: (sp.csr_matrix(v) if sp.issparse(v) else np.asarray(v)) for k, v in d.items() if not k.startswith("__")}


def test_reference_inputs_match_the_generator():
    """The committed .mat is exactly what scripts/make_reference_inputs.py produces from the fixtures."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mri", os.path.join(ROOT, "scripts", "make_reference_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh, stored = mod.build_inputs(), inputs()
    assert set(fresh) == set(stored)
    for k, v in fresh.items():
        a = v.toarray() if sp.issparse(v) else np.atleast_2d(np.asarray(v, dtype=np.float64))
        b = stored[k].toarray() if sp.issparse(stored[k]) else np.atleast_2d(stored[k])
        assert a.shape == b.shape and np.array_equal(a, b), k


def test_generator_script_calls_only_reference_functions():
    src = open(os.path.join(ROOT, "scripts", "make_reference_goldens.m")).read()
    for fn in ("trace_fun_update", "fun_update", "function_multiple_entries", "expmv", "normAm", "lanczos_krylov",
               "mc_trace", "greedy_krylov"):
        assert fn + "(" in src
    assert "addpath(fullfile(refdir, 'functions'))" in src


def _fun_update(P, A, U, B, fun, tol, it, basis):
    """nargout == 4 of functions/fun_update.m:69 is `want_basis` in the oracle and `nargout` in the device package."""
    import inspect
    if "nargout" in inspect.signature(P.fun_update).parameters:
        return P.fun_update(A, U, B, fun, tol, it, 0, nargout=4 if basis else 3)
    return P.fun_update(A, U, B, fun, tol, it, 0, want_basis=basis)


def _check(P, ref, A_of):
    """P: the oracle module or the device package (same function names and argument order)."""
    I = inputs()
    A0 = A_of(I["A0"])
    n0 = I["A0"].shape[0]
    tol0 = float(I["A0_tol"].ravel()[0])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for kind, sgn in (("break", -1.0), ("make", 1.0)):
            E = I["A0_%s_edges" % kind].astype(np.int64)
            for h, (i, j) in enumerate(E):
                U, B = edge_UB(n0, int(i), int(j), sgn)
                x, it, lk = P.trace_fun_update(A0, U, B, tol0, 100, 0, "exp")
                assert it == ref["A0_%s_iter" % kind][h] and bool(lk) == bool(ref["A0_%s_lucky" % kind][h])
                assert abs(x - ref["A0_%s_x" % kind][h]) <= RTOL * abs(ref["A0_%s_x" % kind][h])
        Rome = A_of(I["Rome"])
        for h, (i, j) in enumerate(I["Rome_edges"].astype(np.int64)):
            U, B = edge_UB(I["Rome"].shape[0], int(i), int(j), -1.0)
            x, it, _ = P.trace_fun_update(Rome, U, B, float(I["Rome_tol"].ravel()[0]), 100, 0, "sinh")
            assert it == ref["Rome_sinh_iter"][h] and abs(x - ref["Rome_sinh_x"][h]) <= RTOL * abs(ref["Rome_sinh_x"][h])
        Mx = A_of(I["Mexico"])
        tolM = float(I["Mexico_tol"].ravel()[0])
        Xm, it3, _ = _fun_update(P, Mx, I["Mexico_U"], I["Mexico_B"], "exp", tolM, 100, False)[:3]
        assert it3 == ref["Mexico_lanczos_iter"][0] and Xm.shape[0] == ref["Mexico_lanczos_dim"][0]
        assert abs(np.trace(Xm) - ref["Mexico_lanczos_trace"][0]) <= RTOL * abs(ref["Mexico_lanczos_trace"][0])
        Xm, it4, _, Um = _fun_update(P, Mx, I["Mexico_U"], I["Mexico_B"], "exp", tolM, 100, True)[:4]
        assert it4 == ref["Mexico_arnoldi_iter"][0]
        d = np.einsum("ij,jk,ik->i", Um, Xm, Um)
        assert np.max(np.abs(d - np.asarray(ref["Mexico_arnoldi_update_diag"]))) <= RTOL * np.max(np.abs(ref["Mexico_arnoldi_update_diag"]))
        nrm = P.normest(A0, 1e-6)
        nrm = float(nrm[0] if isinstance(nrm, tuple) else nrm)
        X, itE = P.function_multiple_entries(A0, I["A0_omega"].astype(np.int64), "exp", 1e-10 * np.exp(nrm), 100)
        assert itE == ref["A0_entries_iter"][0]
        assert np.max(np.abs(X - np.asarray(ref["A0_entries"]))) <= RTOL * np.max(np.abs(ref["A0_entries"]))
        f, s, m, mv, mvd, unA = P.expmv(1, A0, I["A0_b"])
        assert [s, m, mv, mvd, unA] == [int(v) for v in ref["A0_expmv_info"]]
        assert np.max(np.abs(f.ravel(order="F") - np.asarray(ref["A0_expmv_f"]))) <= RTOL * np.max(np.abs(ref["A0_expmv_f"]))
        c9, mv9 = P.normAm(A0, 9)
        assert mv9 == ref["A0_normAm9"][1] and abs(c9 - ref["A0_normAm9"][0]) <= 1e-12 * ref["A0_normAm9"][0]
        pr = I["A0_probes"]
        tr, res, itm = P.mc_trace(A0, n0, 1e-3, 60, 1, 0, probes=[(pr[:, :10], pr[:, 10:20]), (pr[:, 20:30], pr[:, 30:40])])
        assert itm == ref["A0_mc_trace"][2] and abs(tr - ref["A0_mc_trace"][0]) <= RTOL * abs(ref["A0_mc_trace"][0])
        edges, rob, _ = P.greedy_krylov(I["A0"], 5, 50, I["A0_centrality"].ravel(), "min", tol0, 100, np.inf, 0, "break")
        assert np.array_equal(np.asarray(edges).ravel(order="F"), np.asarray(ref["A0_greedy_edges"]).astype(np.int64))
        assert abs(rob - ref["A0_greedy_rob"][0]) <= RTOL * abs(ref["A0_greedy_rob"][0])


@pytest.mark.skipif(not os.path.exists(REF), reason="tests/golden/reference_golden.json absent: run "
                    "scripts/make_reference_goldens.m under Octave/MATLAB against the reference (parity unpinned until then)")
def test_oracle_matches_reference_goldens():
    import oracle as O
    _check(O, json.load(open(REF)), lambda A: A)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF), reason="tests/golden/reference_golden.json absent (see above)")
def test_device_matches_reference_goldens():
    import krylov_robustness_b200 as kr
    _check(kr, json.load(open(REF)), lambda A: kr.Matrix(A))
