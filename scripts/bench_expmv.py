#!/usr/bin/env python
"""Config C4 shape: expmv Taylor path on an R-MAT graph with 64 right-hand sides, normAm
degree selection on the device.  Reports ms per Taylor term, matvecs/s and the achieved fraction of the
HBM roofline for the fused Taylor-term kernel (algorithmic bytes 12*nnz + 4*(n+1) + 32*n*q).

  python scripts/bench_expmv.py [--scale 22] [--nnz 67108864] [--q 64] [--norm 8.0] [--device-gen]

--device-gen builds the R-MAT edge list with torch on the GPU (setup only, untimed): the NumPy generator
needs > 6 minutes of host time at the full C4 size (scale 24, 2^28 stored entries); t = norm / lambda_max
is then passed to expmv instead of rescaling the matrix.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rmat_graph_torch(scale, nnz, abcd=(0.57, 0.19, 0.19, 0.05), seed=2):
    """Same construction as krylov_robustness_b200.graphs.rmat_graph (mirror, dedup, fixed-stride thinning)
    with the random bits, the sort/unique and the CSR assembly done by torch on the GPU."""
    import scipy.sparse as sp
    import torch
    dev = torch.device("cuda", 0)
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
    return sp.csr_matrix((np.ones(cols.size), cols, indptr), shape=(n, n))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--nnz", type=int, default=1 << 26)
    ap.add_argument("--q", type=int, default=64)
    ap.add_argument("--norm", type=float, default=8.0, help="spectral radius after scaling (t*lambda_max)")
    ap.add_argument("--device-gen", action="store_true")
    args = ap.parse_args()
    import torch
    import krylov_robustness_b200 as kr
    from krylov_robustness_b200.graphs import rmat_graph, spectral_radius_estimate
    t0 = time.time()
    tscale = 1.0
    if args.device_gen:
        A = rmat_graph_torch(args.scale, args.nnz, seed=2)
        n, nnz = A.shape[0], A.nnz
        gen_s = time.time() - t0
        ctx = kr.Context.default(0)
        M = kr.Matrix(A, ctx)
        lam = float(kr.normest(M, 1e-3)[0])      # symmetric: ||A||_2 = spectral radius (device power iteration)
        tscale = args.norm / lam
    else:
        A = rmat_graph(args.scale, args.nnz, seed=2)
        lam = spectral_radius_estimate(A, 30)
        A = (A * (args.norm / lam)).tocsr()
        n, nnz = A.shape[0], A.nnz
        gen_s = time.time() - t0
        ctx = kr.Context.default(0)
        M = kr.Matrix(A, ctx)
    upload_s = time.time() - t0 - gen_s
    b = np.random.default_rng(3).standard_normal((n, args.q))
    kr.expmv(tscale, M, b[:, :8])                  # warm-up (module load, attributes)
    c0 = ctx.counters()
    ctx.set_timing(True)
    ctx.spmm_time(reset=True)
    t0 = time.perf_counter()
    f, s, m, mv, mvd, unA = kr.expmv(tscale, M, b)
    wall = time.perf_counter() - t0
    ms, launches = ctx.spmm_time(reset=True)
    ctx.set_timing(False)
    terms = mv - mvd
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    bytes_term = 12.0 * nnz + 4.0 * (n + 1) + 32.0 * n * args.q
    # SpMM launches = terms (q columns) + the norm-estimation SpMVs; attribute the measured SpMM time to the terms
    ms_term = ms / max(terms, 1)
    print(json.dumps({"workload": "C4 shape: R-MAT scale %d, nnz %d, q=%d, ||tA||_2~%.1f" % (args.scale, nnz, args.q, args.norm),
                      "n": n, "nnz": nnz, "s": s, "m": m, "mv": mv, "mvd": mvd, "unA": unA,
                      "wall_s_incl_h2d_d2h": wall, "spmm_ms_total": ms, "spmm_launches": launches,
                      "ms_per_taylor_term": ms_term, "matvecs_per_sec": terms * args.q / (ms * 1e-3),
                      "taylor_term_algorithmic_GBs": bytes_term / (ms_term * 1e-3) / 1e9,
                      "frac_of_hbm_peak": bytes_term / (ms_term * 1e-3) / 1e9 / peak, "peak_GBs": peak,
                      "graph_gen_s": gen_s, "matrix_upload_s": upload_s, "t": tscale, "lambda_max": lam, "launches": ctx.counters()["launches"] - c0["launches"]}))


if __name__ == "__main__":
    main()
