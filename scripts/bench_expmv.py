#!/usr/bin/env python
"""Config C4 shape: expmv Taylor path on an R-MAT graph with 64 right-hand sides, normAm
degree selection on the device.  Reports ms per Taylor term, matvecs/s and the achieved fraction of the
HBM roofline for the fused Taylor-term kernel (algorithmic bytes 12*nnz + 4*(n+1) + 32*n*q).

  python scripts/bench_expmv.py [--scale 22] [--nnz 67108864] [--q 64] [--norm 8.0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--nnz", type=int, default=1 << 26)
    ap.add_argument("--q", type=int, default=64)
    ap.add_argument("--norm", type=float, default=8.0, help="spectral radius after scaling (t*lambda_max)")
    args = ap.parse_args()
    import torch
    import krylov_robustness_b200 as kr
    from krylov_robustness_b200.graphs import rmat_graph, spectral_radius_estimate
    t0 = time.time()
    A = rmat_graph(args.scale, args.nnz, seed=2)
    lam = spectral_radius_estimate(A, 30)
    A = (A * (args.norm / lam)).tocsr()
    n, nnz = A.shape[0], A.nnz
    gen_s = time.time() - t0
    ctx = kr.Context.default(0)
    M = kr.Matrix(A, ctx)
    b = np.random.default_rng(3).standard_normal((n, args.q))
    kr.expmv(1, M, b[:, :8])                       # warm-up (module load, attributes)
    c0 = ctx.counters()
    ctx.set_timing(True)
    ctx.spmm_time(reset=True)
    t0 = time.perf_counter()
    f, s, m, mv, mvd, unA = kr.expmv(1, M, b)
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
                      "graph_gen_s": gen_s, "launches": ctx.counters()["launches"] - c0["launches"]}))


if __name__ == "__main__":
    main()
