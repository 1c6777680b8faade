"""The MATLAB/Octave MEX gateway (krylov_robustness_b200/mex/kr_mex.c) compiled against a stub of the MEX API
(tests/mex_stub/) - there is no MATLAB/Octave in the image - and, on the GPU tier, executed with
MATLAB-shaped arguments (CSC sparse A, column-major doubles, 1-based index lists stored as doubles,
func2str strings): the outputs must equal what the Python binding gets from the same C ABI."""
import os
import subprocess
import warnings

import numpy as np
import pytest

from conftest import ROOT, load_graph

STUB = os.path.join(ROOT, "tests", "mex_stub")
PKG = os.path.join(ROOT, "krylov_robustness_b200")


def build_harness(tmp_path):
    exe = str(tmp_path / "mex_harness")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I" + STUB, "-I" + os.path.join(ROOT, "include"),
           os.path.join(STUB, "harness.c"), os.path.join(PKG, "mex", "kr_mex.c"), os.path.join(STUB, "mex_stub.c"),
           "-o", exe, "-L" + PKG, "-l:libkrylov_b200.so", "-Wl,-rpath," + PKG, "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_gateway_compiles_and_links(tmp_path):
    assert os.path.exists(os.path.join(PKG, "libkrylov_b200.so")), "run __graft_entry__.build() first"
    exe = build_harness(tmp_path)
    # every op string a wrapper in matlab/*.m sends is one the gateway dispatches on
    src = open(os.path.join(PKG, "mex", "kr_mex.c")).read()
    import re
    ops = set(re.findall(r"kr_mex\('([a-z_]+)'", "".join(open(os.path.join(PKG, "matlab", f)).read()
                                                            for f in os.listdir(os.path.join(PKG, "matlab")))))
    assert ops, "no kr_mex calls found in matlab/*.m"
    for op in ops:
        assert '"%s"' % op in src, op
    # no GPU here: the harness must fail loudly through mexErrMsgIdAndTxt, not fall back to anything
    if not os.path.exists("/dev/nvidia0"):
        inp = tmp_path / "in.txt"
        inp.write_text("2 2\n0 1 2\n1 0\n1 1\n1\n1 1\n1\n1 2\n1.0 1e-6 10 exp\n")
        r = subprocess.run([exe, str(inp), str(tmp_path / "out.txt")], capture_output=True, text=True)
        assert r.returncode == 3 and "MEX error" in r.stderr


@pytest.mark.gpu
def test_gateway_matches_python_binding(tmp_path):
    import krylov_robustness_b200 as kr
    import oracle as O
    exe = build_harness(tmp_path)
    A = load_graph("oregon_A0")
    n = A.shape[0]
    C = A.tocsc()
    C.sort_indices()
    X = kr.rademacher_host(n, 3, 11)
    c = O.compute_centrality(A, "eig")
    E = O.find_top_missing_edges(A, c, 6, "min")
    nrm, _ = O.normest(A, 1e-2)
    tol = 1e-6 * float(np.exp(nrm))
    with open(tmp_path / "in.txt", "w") as f:
        f.write("%d %d\n" % (n, C.nnz))
        f.write(" ".join(map(str, C.indptr)) + "\n")
        f.write(" ".join(map(str, C.indices)) + "\n")
        f.write(" ".join(repr(float(v)) for v in C.data) + "\n")
        f.write("3\n" + " ".join(repr(float(v)) for v in X.ravel(order="F")) + "\n")
        f.write("%d\n" % len(E) + " ".join(str(float(v)) for v in np.asarray(E, dtype=float).ravel(order="F")) + "\n")
        f.write("1.0 %r 100 exp\n" % tol)
    r = subprocess.run([exe, str(tmp_path / "in.txt"), str(tmp_path / "out.txt")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = {}
    for line in open(tmp_path / "out.txt"):
        p = line.split()
        out[p[0]] = np.array([float(v) for v in p[2:]])
        assert len(out[p[0]]) == int(p[1])
    assert np.array_equal(out["spmm"].reshape(n, 3, order="F"), A @ X)          # 0/1 matrix, +-1 block: exact
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x, it, lucky = kr.trace_fun_update_edges(A, E, 1.0, tol, 100, "exp")
        assert np.array_equal(out["edges_iter"], it) and np.array_equal(out["edges_lucky"], lucky)
        assert np.max(np.abs(out["edges_x"] - x) / np.abs(x)) <= 1e-13
        e, cnt = kr.normest(A, 1e-2)
        assert out["normest"][1] == cnt and abs(out["normest"][0] - e) <= 1e-14 * e
        Xe, ite = kr.function_multiple_entries(A, E, "exp", tol, 100)
        assert out["entries_iter"][0] == ite and np.max(np.abs(out["entries"] - Xe)) <= 1e-13 * np.max(np.abs(Xe))
        f, s, m, mv, mvd, unA = kr.expmv(1, A, X)
        assert tuple(out["expmv_info"]) == (s, m, mv, mvd, unA)
        assert np.max(np.abs(out["expmv_f"].reshape(n, 3, order="F") - f)) <= 1e-13 * np.max(np.abs(f))
        V, H, params, _ = kr.lanczos_krylov(A, X)
        V, H, params, _ = kr.lanczos_krylov(V, H, params)
        assert tuple(out["lanczos_Vdims"]) == V.shape
        assert np.max(np.abs(out["lanczos_H"].reshape(H.shape, order="F") - H)) <= 1e-13 * np.max(np.abs(H))
        # fun_update with the basis (Arnoldi variant), U = unit columns at E[:4, 0], B tridiagonal
        rk = min(4, len(E))
        U = np.zeros((n, rk)); U[E[:rk, 0] - 1, np.arange(rk)] = 1.0
        B = np.zeros((rk, rk))
        for q in range(rk - 1):
            B[q, q + 1] = B[q + 1, q] = 0.1 * (q + 1)
        Xm, itf, lk, Um = kr.fun_update(A, U, B, "exp", tol, 100, 0, nargout=4)
        assert tuple(out["fun_update_info"]) == (itf, float(lk))
        assert np.max(np.abs(out["fun_update_Xm"].reshape(Xm.shape, order="F") - Xm)) <= 1e-12 * np.max(np.abs(Xm))
        assert np.max(np.abs(out["fun_update_Um"].reshape(Um.shape, order="F") - Um)) <= 1e-13
        Xw, dfA = 0.01 * np.arange(1, len(E) + 1), np.full(len(E), 0.5)
        f2, gr2 = kr.fun_and_grad_krylov_fun(Xw, A, E, "sinh", "cosh", dfA, 1e-8, 100)
        assert abs(out["fg_f"][0] - f2) <= 1e-12 * abs(f2) and np.max(np.abs(out["fg_gr"] - gr2)) <= 1e-12 * np.max(np.abs(gr2))
        import scipy.sparse as sp
        Hes = kr.hessianfcn_exp(np.zeros(len(E)), A, E, 1e-10, 100)
        assert np.max(np.abs(out["hessian"].reshape(Hes.shape, order="F") - Hes)) <= 1e-12 * np.max(np.abs(Hes))
        probes = out["mc_probes"].reshape(n, 20, order="F")
        tr, res, itm = kr.mc_trace(A, n, 1e-3, 30, 1, 0, probes=[(probes[:, :10], probes[:, 10:])])
        assert out["mc_trace"][2] == itm and abs(out["mc_trace"][0] - tr) <= 1e-12 * abs(tr)
    # same mxArray, doubled values: a cache keyed on (data pointer, nnz) would have served the old matrix
    assert np.array_equal(out["spmm_doubled_values"].reshape(n, 3, order="F"), 2.0 * (A @ X))
    # explicit handle + in-place edge insertion on the device copy == scoring on the edited matrix
    A2 = sp.lil_matrix(A)
    A2[E[0, 0] - 1, E[0, 1] - 1] = 1.0
    A2[E[0, 1] - 1, E[0, 0] - 1] = 1.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x2, _, _ = kr.trace_fun_update_edges(A2.tocsr(), E, 1.0, tol, 100, "exp")
    assert np.max(np.abs(out["edges_x_after_insert"] - x2) / np.abs(x2)) <= 1e-12
    # one greedy 'make' round through the gateway (the call matlab/krylov_miobi.m makes): 1-based winner, its value, count
    bsel, vsel = kr.select_candidate(x2, "make")
    assert out["greedy_round_make"][0] == bsel + 1 and abs(out["greedy_round_make"][1] - vsel) <= 1e-12 * abs(vsel)
    assert out["greedy_round_make"][2] == len(E)
