"""Oracle: greedy edge break/make drivers and candidate generation.  TEST INFRASTRUCTURE ONLY.

Restates functions/krylov_miobi.m, greedy_krylov.m, find_top_edges.m, find_top_missing_edges.m,
edge2low_rank.m and the 'eig' / 'deg' branches of compute_centrality.m.  Edge lists are k x 2
integer arrays with the reference's 1-based node numbers.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .updates import trace_fun_update


def _issymmetric(A):
    A = sp.csr_matrix(A)
    return (A != A.T).nnz == 0


def edge2low_rank(E, n, sign=-1.0):
    """[U,B] = edge2low_rank(E,n)  (edge2low_rank.m:1-13; B entries -1 there and in the break
    scripts, +1 in Tests/test_unweighted_make.m:171-183 - hence the ``sign`` argument)."""
    E = np.atleast_2d(np.asarray(E)).astype(np.int64)
    ut = np.unique(E.ravel())
    pos = {int(a): i for i, a in enumerate(ut)}
    U = sp.csr_matrix((np.ones(ut.size), (ut - 1, np.arange(ut.size))), shape=(n, ut.size))
    B = np.zeros((ut.size, ut.size))
    for a, b in E:
        B[pos[int(a)], pos[int(b)]] = sign
        B[pos[int(b)], pos[int(a)]] = sign
    return U, B


def compute_centrality(A, kind="eig"):
    """c = compute_centrality(A,type)  (compute_centrality.m:9-10 'exp', :15-17 'eig', :18-19 'deg', :20-26 'pr';
    'res' (:11-14) uses an undefined variable in the reference and is not restated)."""
    if kind == "deg":
        return np.asarray(A.sum(axis=0)).ravel()
    n = A.shape[0]
    A = sp.csr_matrix(A).astype(np.float64)
    if kind == "exp":
        import scipy.linalg as sla
        return np.diag(sla.expm(A.toarray())).copy()
    if kind == "pr":
        alpha = 0.85
        dinv = 1.0 / np.asarray(A.sum(axis=0)).ravel()
        op = spla.LinearOperator((n, n), dtype=np.float64,
                                 matvec=lambda x: alpha * (A @ (dinv * np.ravel(x))) + (1 - alpha) * np.sum(x) / n * np.ones(n))
        _, u = spla.eigs(op, k=1, which="LM", v0=np.ones(n), tol=0)
        return np.abs(u[:, 0])
    _, u = spla.eigsh(A, k=1, which="LM", v0=np.ones(n), tol=1e-12)
    return np.abs(u[:, 0])


def _tril_edges(A):
    # [I,J] = find(tril(A,-1)) : column-major order, 1-based
    L = sp.tril(sp.csc_matrix(A), k=-1, format="csc")
    L.sort_indices()
    L.eliminate_zeros()
    J = np.repeat(np.arange(L.shape[1]), np.diff(L.indptr)) + 1
    I = L.indices + 1
    return I.astype(np.int64), J.astype(np.int64)


def find_top_edges(A, centrality, num, order="mult"):
    """E = find_top_edges(A,centrality,num,order)  (find_top_edges.m:1-40)."""
    centrality = np.asarray(centrality, dtype=np.float64).ravel()
    I, J = _tril_edges(A)
    E = np.stack([I, J], axis=1)
    if order == "mult":
        c = centrality[I - 1] * centrality[J - 1]
        ind = np.argsort(-c, kind="stable")
        return E[ind[:num]]
    if order == "min":
        sc = np.sort(centrality)[::-1]
        # find(sc == value, 1): first position in the descending list = 1 + #entries strictly greater
        asc = sc[::-1]
        rank = lambda v: sc.size - np.searchsorted(asc, v, side="right") + 1
        c1 = rank(centrality[I - 1])
        c2 = rank(centrality[J - 1])
        mn = np.minimum(c1, c2).astype(np.float64)
        mx = np.maximum(c1, c2).astype(np.float64)
        scores = mx * (mx - 1) / 2 + mn
        ind = np.argsort(scores, kind="stable")
        return E[ind[:num]]
    raise ValueError(order)


def find_top_missing_edges(A, centrality, num, order="min"):
    """E = find_top_missing_edges(A,centrality,num,order)  (find_top_missing_edges.m:1-67)."""
    centrality = np.asarray(centrality, dtype=np.float64).ravel()
    indC = np.argsort(-centrality, kind="stable")           # 0-based node ids, descending centrality
    Ac = sp.csc_matrix(A)
    if order == "min":                                      # :55-65
        rows = []
        j = 2
        total = 0
        while total < num:
            col = Ac[:, indC[j - 1]].toarray().ravel()
            cand = indC[:j - 1]
            ind = cand[col[cand] == 0]
            rows.append(np.stack([ind + 1, np.full(ind.size, indC[j - 1] + 1)], axis=1))
            total += ind.size
            j += 1
        E = np.concatenate(rows, axis=0)
        return E[:num].astype(np.int64)
    if order == "mult":                                     # :20-54 (main path)
        sc = centrality[indC]
        min_N, length = 2, 0
        while length < num:
            col = Ac[:, indC[min_N - 1]].toarray().ravel()
            length += int(np.sum(col[indC[:min_N - 1]] == 0))
            min_N += 1
        min_N -= 1
        N = int(np.sum(sc[0] * sc > sc[min_N - 1] ** 2))
        S = np.triu(np.outer(sc[:N], sc[:N]))
        ind = np.argsort(-S.ravel(order="F"), kind="stable")
        Ii, Jj = np.unravel_index(ind, S.shape, order="F")
        Ad = Ac[indC[:N], :][:, indC[:N]].toarray()
        out = []
        for a, b in zip(Ii, Jj):
            if len(out) >= num:
                break
            if a != b and Ad[a, b] == 0:
                out.append((indC[a] + 1, indC[b] + 1))
        return np.asarray(out, dtype=np.int64)
    raise ValueError(order)


def _set_edge(A, i, j, val):
    A = A.tolil()
    A[i - 1, j - 1] = val
    A[j - 1, i - 1] = val
    A = A.tocsr()
    A.eliminate_zeros()
    return A


def krylov_miobi(A, k, E=None, tol=1e-12, it=None, poles=np.inf, debug=0, miobi="break", rescale=1.0,
                 scorer=None):
    """[edges, rob, A_new] = krylov_miobi(A,k,E,tol,it,poles,debug,miobi,rescale)  (krylov_miobi.m:1-142).

    ``scorer(A, E, sign, rescale, tol, it)`` (optional) replaces the inner candidate loop
    (:76-99) and must return the vector of trace_fun_update values; it exists so tests can run
    the same driver logic around the device scorer."""
    if not _issymmetric(A):
        raise ValueError("KRYLOV_MIOBI:: Adjacency matrix should be symmetric")
    A = sp.csr_matrix(A).astype(np.float64)
    n = A.shape[0]
    if it is None:
        it = min(100, n)
    if E is None or len(E) == 0:
        Ac = sp.coo_matrix(A)
        keep = Ac.row >= Ac.col
        order = np.lexsort((Ac.row[keep], Ac.col[keep]))
        E = np.stack([Ac.row[keep][order] + 1, Ac.col[keep][order] + 1], axis=1)
    E = np.atleast_2d(np.asarray(E)).astype(np.int64)
    if miobi not in ("break", "make"):
        raise ValueError("KRYLOV_MIOBI:: not supported option for miobi")
    if miobi == "break" and A.nnz < 2 * k:
        raise ValueError("KRYLOV_MIOBI:: edges to be removed are more than edges in the network")
    sign = -1.0 if miobi == "break" else 1.0
    rob = 0.0
    edges = np.zeros((0, 2), dtype=np.int64)
    nE = E.shape[0]
    for _ in range(min(k, nE)):
        if scorer is not None:
            vals = np.asarray(scorer(A, E, sign, rescale, tol, it))
        else:
            vals = np.zeros(nE)
            for h in range(nE):                             # :76-99
                i, j = int(E[h, 0]), int(E[h, 1])
                if i != j:
                    B = sign * np.array([[0.0, 1.0], [1.0, 0.0]]) / rescale
                    U = np.zeros((n, 2))
                    U[i - 1, 0] = 1.0
                    U[j - 1, 1] = 1.0
                else:
                    B = np.array([[sign]])
                    U = np.zeros((n, 1))
                    U[i - 1, 0] = 1.0
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    vals[h] = trace_fun_update(A, U, B, tol, it, debug)[0]
        best, bestval = -1, (np.inf if miobi == "break" else -np.inf)
        for h in range(nE):                                 # :112-124 strict compare, first wins
            if (miobi == "break" and vals[h] < bestval) or (miobi == "make" and vals[h] > bestval):
                best, bestval = h, vals[h]
        chosen = E[best].copy()
        E = np.delete(E, best, axis=0)                      # :127
        nE -= 1
        A = _set_edge(A, int(chosen[0]), int(chosen[1]), 0.0 if miobi == "break" else 1.0)
        edges = np.vstack([edges, chosen[None, :]])
        rob += bestval
    return edges, rob, A


def greedy_krylov(A, k, Q=0, centrality=None, order="mult", tol=1e-12, it=None, poles=np.inf, debug=0,
                  miobi="break", rescale=1.0, scorer=None):
    """[edges, rob_variation, A_new] = greedy_krylov(A,k,Q,centrality,order,tol,it,poles,debug,miobi,rescale)
    (greedy_krylov.m:1-97)."""
    if not _issymmetric(A):
        raise ValueError("GREEDY_KRYLOV:: Adjacency matrix should be symmetric")
    A = sp.csr_matrix(A).astype(np.float64)
    if it is None:
        it = min(100, A.shape[0])
    if not Q:
        Q = int(np.asarray(A.sum(axis=0)).max())
    if miobi == "break" and A.nnz < 2 * k:
        raise ValueError("GREEDY_KRYLOV:: edges to be removed are more than edges in the network")
    rob_variation = 0.0
    edges = np.zeros((0, 2), dtype=np.int64)
    top_edges = None
    tmp_edges = None
    for j in range(1, k + 1):
        if j == 1:
            if miobi == "make":
                top_edges = find_top_missing_edges(A, centrality, Q + k, order)
            else:
                top_edges = find_top_edges(A, centrality, Q + k, order)
        else:                                               # :84-86
            match = np.all(top_edges == tmp_edges, axis=1)
            top_edges = top_edges[~match]
        E = top_edges[:Q]
        tmp_edges, tmp_rob, A = krylov_miobi(A, 1, E, tol, it, poles, debug, miobi, rescale, scorer=scorer)
        edges = np.vstack([edges, tmp_edges])
        rob_variation += tmp_rob
    return edges, rob_variation, A
