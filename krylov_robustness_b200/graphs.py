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


def rmat_graph_device(scale, nnz, abcd=(0.57, 0.19, 0.19, 0.05), seed=2, device=0):
    """Same construction as rmat_graph (mirror, dedup, fixed-stride thinning) with the random bits, the
    sort / unique and the CSR assembly done by torch on the GPU: the NumPy generator needs > 6 minutes of host time
    at the full C4 size (scale 24, 2^28 stored entries).  Setup only, never timed.  Returns (indptr, indices) as
    int64 NumPy arrays of the symmetric 0/1 pattern."""
    import torch
    dev = torch.device("cuda", device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n, m = 1 << scale, nnz // 2
    draw = int(m * 1.25) + 16
    a, b, c, _ = abcd
    i = torch.zeros(draw, dtype=torch.int64, device=dev)
    j = torch.zeros(draw, dtype=torch.int64, device=dev)
    for _ in range(scale):
        r = torch.rand(draw, generator=g, device=dev, dtype=torch.float64)
        right = ((r >= a) & (r < a + b)) | (r >= a + b + c)
        down = r >= a + b
        i = (i << 1) | down.to(torch.int64)
        j = (j << 1) | right.to(torch.int64)
    del r, right, down
    keep = i != j
    lo, hi = torch.minimum(i, j)[keep], torch.maximum(i, j)[keep]
    del i, j, keep
    key = torch.unique(lo * n + hi)
    del lo, hi
    if key.numel() > m:
        sel = torch.linspace(0, key.numel() - 1, m, dtype=torch.float64, device=dev).to(torch.int64)
        key = key[sel]
    lo, hi = key // n, key % n
    full = torch.sort(torch.cat([lo * n + hi, hi * n + lo])).values     # row-major order of both triangles
    del lo, hi, key
    rows, cols = full // n, full % n
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    indptr, cols = indptr.cpu().numpy(), cols.cpu().numpy()
    del full, rows
    torch.cuda.empty_cache()
    return indptr, cols


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
