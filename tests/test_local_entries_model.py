"""CPU tier: the ALGORITHM of csrc/entries_local.cuh restated in NumPy and checked against the oracle.

The device kernel runs one single-vector Arnoldi space of function_multiple_entries (functions/function_multiple_entries.m:86-164,
arnoldi_krylov.m:104-106) on the ball around its start node: local numbering in BFS order (a new level sorted by global id),
basis vector l stored with |ball(l)| entries, CGS2 + third pass, projected solve x = f(T) e1 on the tridiagonal band of H by a
scaled Taylor evaluation of exp(+-T) e1, lag-3 stop.  This file pins that formulation itself - that restricting the process
to the ball, dropping the off-band rounding noise of H and replacing the eigen-decomposition by the Taylor scheme reproduces
the reference's entries and iteration counts - independently of the CUDA code (whose parity tests are in test_gpu_krylov.py).
"""
import math
import warnings

import numpy as np
import pytest
import scipy.linalg as sl
import scipy.sparse as sp

import oracle as O
from conftest import load_graph

FUN = {"exp": np.exp, "sinh": np.sinh, "cosh": np.cosh}


def taylor_fx(d, e, fun):
    """x = f(T) e1, T = tridiag(e, d, e): warp_tridiag_fx_taylor (s = ceil(||T||_inf) stages of degree 20)."""
    n = len(d)
    T = np.diag(d) + (np.diag(e, 1) + np.diag(e, -1) if n > 1 else 0.0)
    rho = float(np.max(np.abs(T).sum(1)))
    ns = math.ceil(rho) if rho > 1.0 else 1
    e1 = np.zeros(n)
    e1[0] = 1.0
    if ns == 1:
        t, ev, od = e1.copy(), e1.copy(), np.zeros(n)
        for k in range(1, 21):
            t = (T @ t) * (1.0 / k)
            if k & 1:
                od = od + t
            else:
                ev = ev + t
        return {"exp": ev + od, "sinh": od, "cosh": ev}[fun]
    yp, ym = e1.copy(), e1.copy()
    for _ in range(ns):
        tp, tm = yp.copy(), ym.copy()
        for k in range(1, 21):
            ck = (1.0 / ns) / k
            tp = (T @ tp) * ck
            yp = yp + tp
            tm = -(T @ tm) * ck
            ym = ym + tm
    return {"exp": yp, "sinh": 0.5 * (yp - ym), "cosh": 0.5 * (yp + ym)}[fun]


def local_space(A, h, j2s, fun, tol, it):
    """One space K(A, e_h) on the ball around h; returns the entries f(A)(h, j2), the steps taken, |ball|, arena use."""
    ip, ix, dv = A.indptr, A.indices, A.data
    L, loc, lens, V, lev_begin = [h], {h: 0}, [1], [np.array([1.0])], 0
    H = np.zeros((it + 2, it + 1))
    ring = {}
    x, jj = None, 0
    for j in range(it):
        m_old = len(L)
        new = set()
        for idx in range(lev_begin, m_old):
            u = L[idx]
            for g in ix[ip[u]:ip[u + 1]]:
                if g not in loc:
                    new.add(int(g))
        for g in sorted(new):
            loc[g] = len(L)
            L.append(g)
        m = len(L)
        lev_begin = m_old
        lens.append(m)
        vj, lj = V[j], lens[j]
        w = np.zeros(m)
        for r in range(m):
            u = L[r]
            s = 0.0
            for p in range(ip[u], ip[u + 1]):
                li = loc.get(int(ix[p]), -1)
                if 0 <= li < lj:
                    s += dv[p] * vj[li]
            w[r] = s
        rr = 0.0
        for ps in range(3):
            hc = [float(V[l] @ w[:lens[l]]) for l in range(j + 1)]
            for l in range(j + 1):
                w[:lens[l]] -= hc[l] * V[l]
                H[l, j] = hc[l] if ps == 0 else H[l, j] + hc[l] * (rr if ps == 2 else 1.0)
            if ps == 1:
                rr = math.sqrt(float(w @ w))
                w *= 1.0 / rr if rr > 0.0 else 0.0
                H[j + 1, j] = rr
        V.append(w)
        jj = j + 1
        d = np.array([H[i, i] for i in range(jj)])
        e = np.array([0.5 * (H[i + 1, i] + H[i, i + 1]) for i in range(jj - 1)])
        x = taylor_fx(d, e, fun)
        ring[jj] = x
        if jj > 3:
            old = np.zeros(jj)
            old[:jj - 3] = ring[jj - 3]
            if not (np.linalg.norm(x - old) > tol):
                break
    out = []
    for j2 in j2s:
        li = loc.get(int(j2), -1)
        out.append(0.0 if li < 0 else sum(V[l][li] * x[l] for l in range(jj) if li < lens[l]))
    return np.array(out), jj, len(L), sum(lens[:jj + 1])


@pytest.mark.parametrize("fun", ["exp", "sinh", "cosh"])
def test_taylor_solve_matches_expm(fun):
    rng = np.random.default_rng(0)
    for _ in range(60):
        n = int(rng.integers(1, 25))
        scale = float(rng.choice([1e-3, 0.05, 0.8, 3.5, 12.0, 40.0]))
        d = rng.standard_normal(n) * 0.1 * scale
        e = np.abs(rng.standard_normal(max(n - 1, 0))) * scale / 2.5
        T = np.diag(d) + (np.diag(e, 1) + np.diag(e, -1) if n > 1 else 0.0)
        Ep, Em = sl.expm(T)[:, 0], sl.expm(-T)[:, 0]
        ref = {"exp": Ep, "sinh": 0.5 * (Ep - Em), "cosh": 0.5 * (Ep + Em)}[fun]
        x = taylor_fx(d, e, fun)
        # sinh by the expm pair itself cancels for small ||T||: compare on the scale of exp there
        denom = max(np.max(np.abs(ref)), np.max(np.abs(Ep)) * 1e-3 if fun == "sinh" else 0.0)
        rho = float(np.max(np.abs(T).sum(1)))
        # both sides lose digits in proportion to ||T|| (squarings in expm, stages here)
        assert np.max(np.abs(x - ref)) <= 5e-13 * max(1.0, rho) * denom


def test_local_space_matches_oracle_on_road_network():
    """Config C2's graph: entries over edges, one space per pair, against the oracle's function_multiple_entries."""
    A = load_graph("transport_Vermont").astype(np.float64)
    A = (A / A.max()).tocsr()
    Lc = sp.tril(A, -1).tocoo()
    sel = np.random.default_rng(17).choice(Lc.nnz, 5, replace=False)
    om = np.stack([Lc.row[sel] + 1, Lc.col[sel] + 1], 1)
    om = np.vstack([om, [om[0, 0], om[0, 0]], [om[1, 0], 77]])            # a diagonal entry, a far-away second index
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nrm, _ = O.normest(A, 1e-2)
        tol = 1e-8 * float(np.cosh(nrm))
        for p in range(om.shape[0]):
            oX, oit = O.function_multiple_entries(A, om[p:p + 1], "cosh", tol, 100)
            x, steps, ball, arena = local_space(A, int(om[p, 0] - 1), [int(om[p, 1] - 1)], "cosh", tol, 24)
            assert steps == oit
            assert abs(x[0] - oX[0]) <= 1e-10 * max(1.0, abs(oX[0]))
            assert ball <= 384 and arena <= 2048                        # the first budget tier of the kernel holds them


def test_local_space_exhausted_component_and_self_loop():
    """Lucky breakdown (the component is exhausted, the new vector is exactly zero) and a diagonal entry of A."""
    ring = sp.diags([np.ones(7), np.ones(7)], [-1, 1], shape=(8, 8)).tolil()
    ring[0, 7] = ring[7, 0] = 1.0
    ring[2, 2] = 0.5
    path = sp.diags([np.ones(3), np.ones(3)], [-1, 1], shape=(4, 4))
    A = sp.block_diag([sp.csr_matrix(ring), sp.csr_matrix(path), sp.csr_matrix((1, 1))]).tocsr().astype(np.float64)
    om = np.array([[1, 1], [3, 3], [2, 6], [9, 12], [11, 10], [13, 13], [1, 10]])
    F = sl.expm(A.toarray())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for p in range(om.shape[0]):
            x, steps, _, _ = local_space(A, int(om[p, 0] - 1), [int(om[p, 1] - 1)], "exp", 1e-12, 20)
            oX, oit = O.function_multiple_entries(A, om[p:p + 1], "exp", 1e-12, 20)
            assert np.isfinite(x[0])
            assert abs(x[0] - F[om[p, 0] - 1, om[p, 1] - 1]) <= 1e-10
            assert abs(x[0] - oX[0]) <= 1e-10
            assert steps == oit
