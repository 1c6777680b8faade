"""Seeded synthetic graph generators for the bench / full-size property tests.

Host-side NumPy only (setup, never timed): a Chung-Lu power-law graph (config C3 / C5) and an
R-MAT graph (config C4).  Both return a symmetric 0/1 CSR matrix with zero diagonal.
"""
import numpy as np
import scipy.sparse as sp


def _symmetric_from_edges(n, i, j, target_edges=None):
    keep = i != j
    i, j = i[keep], j[keep]
    lo = np.minimum(i, j).astype(np.int64)
    hi = np.maximum(i, j).astype(np.int64)
    key = np.unique(lo * n + hi)
    if target_edges is not None and key.size > target_edges:
        # deterministic thinning: keep a fixed stride subset
        sel = np.linspace(0, key.size - 1, target_edges).astype(np.int64)
        key = key[sel]
    lo, hi = key // n, key % n
    rows = np.concatenate([lo, hi])
    cols = np.concatenate([hi, lo])
    A = sp.csr_matrix((np.ones(rows.size), (rows, cols)), shape=(n, n))
    A.sort_indices()
    return A


def power_law_graph(n=1_000_000, nnz=20_000_000, gamma=2.2, seed=20260310):
    """Chung-Lu graph with expected degrees w_i ~ (i + i0)^(-1/(gamma-1)), capped at sqrt(nnz)
    (structural cut-off), target ``nnz`` stored entries (both triangles)."""
    rng = np.random.default_rng(seed)
    m = nnz // 2
    a = 1.0 / (gamma - 1.0)
    i0 = 1.0
    for _ in range(60):                      # pick i0 so that the largest expected degree ~ sqrt(nnz)
        w = (np.arange(n) + i0) ** (-a)
        w *= nnz / w.sum()
        if w[0] <= np.sqrt(nnz):
            break
        i0 *= 1.5
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    draw = int(m * 1.12) + 16
    i = np.searchsorted(cdf, rng.random(draw)).astype(np.int64)
    j = np.searchsorted(cdf, rng.random(draw)).astype(np.int64)
    np.minimum(i, n - 1, out=i)
    np.minimum(j, n - 1, out=j)
    # scatter node ids so that degree is not monotone in the index (no free locality)
    perm = rng.permutation(n)
    return _symmetric_from_edges(n, perm[i], perm[j], target_edges=m)


def rmat_graph(scale=24, nnz=1 << 28, abcd=(0.57, 0.19, 0.19, 0.05), seed=2):
    """R-MAT graph on 2^scale nodes, mirrored and deduplicated to ~nnz stored entries."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    m = nnz // 2
    draw = int(m * 1.25) + 16
    a, b, c, _ = abcd
    i = np.zeros(draw, dtype=np.int64)
    j = np.zeros(draw, dtype=np.int64)
    for _ in range(scale):
        r = rng.random(draw)
        right = (r >= a) & (r < a + b) | (r >= a + b + c)
        down = r >= a + b
        i = (i << 1) | down
        j = (j << 1) | right
    return _symmetric_from_edges(n, i, j, target_edges=m)


def spectral_radius_estimate(A, iters=30):
    """Power-iteration estimate of lambda_max used to scale A so that exp stays in range
    (SURVEY.md 7.2.7)."""
    x = np.ones(A.shape[0]) / np.sqrt(A.shape[0])
    lam = 0.0
    for _ in range(iters):
        y = A @ x
        lam = float(np.linalg.norm(y))
        x = y / lam
    return lam
