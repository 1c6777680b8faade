"""The drop-in boundary exercised in the REFERENCE'S OWN LANGUAGE.

scripts/make_reference_goldens.m - the script whose run over the reference's functions/*.m produced
tests/golden/reference_golden.json - is executed a second time, unchanged, with the MATLAB path pointing at this
repository's wrappers (krylov_robustness_b200/matlab/*.m) instead of the reference's functions: the interpreter
(oracle/mlab) runs the wrapper .m files, `kr_mex(...)` is the real MEX gateway (mex/kr_mex.c, compiled against the stub
MEX runtime and loaded through tests/mex_stub/bridge.py), and the gateway calls the C ABI of libkrylov_b200.so on the
GPU.  Same inputs, same call sites, same output keys: the numbers must be the reference's (1e-10, counts equal)."""
import io
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

from oracle.mlab import Interpreter, MatlabError

import sys
sys.path.insert(0, os.path.join(ROOT, "tests", "mex_stub"))
import bridge  # noqa: E402

WRAPPERS = os.path.join(ROOT, "krylov_robustness_b200", "matlab")
SCRIPT = os.path.join(ROOT, "scripts", "make_reference_goldens.m")

INT_KEYS = {"A0_break_iter", "A0_break_lucky", "A0_make_iter", "A0_make_lucky", "Rome_sinh_iter", "Rome_cosh_iter",
            "Mexico_arnoldi_iter", "Mexico_arnoldi_dim", "Mexico_arnoldi_sinh_iter", "A0_entries_iter", "A0_expmv_info",
            "A0_expmv_half_full_term_info", "A0_loops_expmv_info", "A0_std_info", "A0_arnoldi_dims", "A0_edge_set_rk",
            "A0_miobi_rescale_edges", "A0_miobi_rescale_nnz"}
# value followed by counts / flags in one vector
MIXED = {"A0_selfloop": 1, "A0_dense_branch": 1, "A0_edge_set": 1, "A0_normAm9": 1, "A0_loops_normAm5": 1, "A0_mc_trace": 2}
SKIP = {
    "Mexico_lanczos_form_raises",   # the reference's three-output fun_update raises (fun_update.m:137); the wrapper's works
    "Mexico_lanczos_trace", "Mexico_lanczos_iter", "Mexico_lanczos_dim",
    "A0_normest",                   # MATLAB's own normest m-file in both runs, not a wrapper
}


def host(tmp_path):
    so = bridge.build_host(tmp_path)
    H = bridge.MexHost(so)
    I = Interpreter(stdout=io.StringIO())
    H.install(I)
    return H, I


def test_wrappers_parse_and_the_gateway_refuses_to_run_without_a_gpu(tmp_path):
    """CPU tier: every wrapper is valid MATLAB for the interpreter; the gateway library builds, loads and reports
    errors back to the MATLAB level; without a CUDA device a wrapper call FAILS (no CPU fallback behind the boundary)."""
    import glob
    from oracle.mlab.parser import parse_source
    files = sorted(glob.glob(os.path.join(WRAPPERS, "*.m")))
    assert len(files) >= 16
    for f in files:
        body, funs = parse_source(open(f).read(), f)
        assert funs and funs[0].name == os.path.splitext(os.path.basename(f))[0]
    H, I = host(tmp_path)
    I.addpath(WRAPPERS)
    with pytest.raises(MatlabError):
        I.call("kr_mex", "no_such_operation", nargout=1)
    # marshalling round trips (interpreter value -> mxArray -> interpreter value)
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    S = sp.random(40, 30, 0.1, format="csc", random_state=1)
    D = rng.standard_normal((7, 3))
    for v in (S, D, D[:, :1].T, "sinh", np.array([[True]]), bridge.UInt64(2**40 + 5), np.zeros((0, 0))):
        p = H.to_mx(v)
        back = H.from_mx(p)
        H.L.mxDestroyArray(p)
        if sp.issparse(v):
            assert (back != v).nnz == 0 and back.shape == v.shape
        elif isinstance(v, bridge.UInt64):
            assert back.v == v.v
        elif isinstance(v, str):
            assert back == v
        else:
            assert back.shape == np.atleast_2d(v).shape and np.array_equal(back, np.atleast_2d(v))
    import torch
    if not torch.cuda.is_available():
        import scipy.sparse as sp
        A = sp.identity(200, format="csc") * 0.5
        U = np.zeros((200, 2))
        U[0, 0] = U[1, 1] = 1.0
        with pytest.raises(MatlabError):
            I.call("trace_fun_update", A, U, np.array([[0.0, 1.0], [1.0, 0.0]]), 1e-8, 50, 0, nargout=3)


@pytest.mark.gpu
def test_generator_script_through_the_wrappers_reproduces_the_reference(tmp_path):
    H, I = host(tmp_path)
    out_json = str(tmp_path / "dropin.json")
    I.run_script(SCRIPT, {"dropin_dir": WRAPPERS, "golden_path": out_json})
    got = json.load(open(out_json))
    ref = json.load(open(os.path.join(GOLDEN, "reference_golden.json")))
    # every call went through the gateway, and the wrappers are what ran (no reference file is on the path)
    assert H.calls.get("trace_fun_update", 0) >= 25 and H.calls.get("krylov_start", 0) >= 2
    for op in ("fun_update", "function_multiple_entries", "expmv", "select_taylor_degree", "normAm", "mc_trace",
               "greedy_round", "matrix_set_edges", "fun_and_grad", "hessian", "krylov_extend"):
        assert H.calls.get(op, 0) >= 1, op
    assert all(os.path.dirname(f) in (WRAPPERS, os.path.dirname(SCRIPT)) or "mlab_" in f for f in I.cache), list(I.cache)
    bad, n = [], 0
    for k, want in ref.items():
        if k in SKIP or k not in got:
            continue
        w, g = np.asarray(want, dtype=np.float64), np.asarray(got[k], dtype=np.float64)
        n += 1
        if w.shape != g.shape:
            bad.append("%s: shape %s vs %s" % (k, g.shape, w.shape))
            continue
        if k in INT_KEYS:
            if not np.array_equal(w, g):
                bad.append("%s: %s != %s" % (k, g.tolist()[:10], w.tolist()[:10]))
            continue
        if k == "A0_trace_exp":          # the wrapper draws all 34 probe pairs up front; the value is what counts
            w, g = w[:1], g[:1]
        if k == "Mexico_fg30_sinh":      # objective from a ~55-column Lanczos block: see tests/test_reference_goldens.py
            if not abs(g[0] - w[0]) <= 1e-3 * abs(w[0]):
                bad.append("%s objective: %r vs %r" % (k, g[0], w[0]))
            w, g = w[1:], g[1:]
        if k in MIXED:
            nv = MIXED[k]
            if not np.array_equal(w[nv:], g[nv:]):
                bad.append("%s counts: %s != %s" % (k, g[nv:].tolist(), w[nv:].tolist()))
            w, g = w[:nv], g[:nv]
        scale = np.max(np.abs(w)) if w.size else 1.0
        err = np.max(np.abs(w - g)) / (scale if scale > 0 else 1.0) if w.size else 0.0
        if not err <= 1e-10:
            bad.append("%s: rel err %.3e" % (k, err))
    # the operator plug-in point (lanczos_krylov.m:32,78-79): kr_operator(A) through the L1 wrappers gives the same
    # projections as the matrix, and its `multiply` handle is the device SpMM
    import scipy.io as sio
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        inp = sio.loadmat(os.path.join(GOLDEN, "reference_inputs.mat"))
    w = np.asarray(ref["A0_lanczos_ritz"])
    g = np.asarray(got["A0_lanczos_ritz_operator_struct"])
    assert g.shape == w.shape and np.max(np.abs(g - w)) <= 1e-10 * np.max(np.abs(w))
    y = (inp["A0"] @ inp["A0_b"]).ravel(order="F")
    assert np.max(np.abs(np.asarray(got["A0_operator_multiply"]) - y)) <= 1e-13 * np.max(np.abs(y)) * 50
    assert len(got["A0_arnoldi_first_block_operator_struct"]) == 3 and H.calls.get("spmm", 0) >= 1
    assert n >= 40, n
    assert not bad, "%d of %d keys differ from the reference's outputs:\n  %s" % (len(bad), n, "\n  ".join(bad))
