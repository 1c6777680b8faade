#!/usr/bin/env python
"""Config C5 shape: score candidate edges (rank-2 block Lanczos per candidate) on the C3 power-law graph.
  python scripts/bench_edges.py [--ncand 1024] [--n 1000000] [--nnz 20000000]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncand", type=int, default=1024)
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--nnz", type=int, default=20_000_000)
    ap.add_argument("--tolfac", type=float, default=1e-6)
    args = ap.parse_args()
    import krylov_robustness_b200 as kr
    from krylov_robustness_b200.graphs import power_law_graph, spectral_radius_estimate
    A = power_law_graph(args.n, args.nnz, 2.2, 20260310)
    lam = spectral_radius_estimate(A, 30)
    A = (A * (1.0 / lam)).tocsr()
    deg = np.diff(A.indptr)
    top = np.argsort(-deg, kind="stable")[:max(64, int(2.2 * np.sqrt(2 * args.ncand)))]
    sub = A[top][:, top].toarray()
    ii, jj = np.where(np.triu(sub == 0, 1))
    E = np.stack([top[jj] + 1, top[ii] + 1], 1)[:args.ncand].astype(np.int64)
    ctx = kr.Context.default(0)
    M = kr.Matrix(A, ctx)
    tol = args.tolfac * float(np.exp(1.0))
    kr.trace_fun_update_edges(M, E[:8], 1.0 / lam, tol, 100, "exp")
    ctx.set_timing(True); ctx.spmm_time(reset=True)
    c0 = ctx.counters()
    t0 = time.perf_counter()
    x, it, lucky = kr.trace_fun_update_edges(M, E, 1.0 / lam, tol, 100, "exp")
    dt = time.perf_counter() - t0
    ms, nl = ctx.spmm_time(reset=True)
    c1 = ctx.counters()
    steps = int(it.max())
    n = A.shape[0]
    step_bytes = 96.0 * n * 2 * E.shape[0]            # SpMM dense side 32 + two update passes 64 B per row per column
    print(json.dumps({"candidates": int(E.shape[0]), "seconds": dt, "edges_per_sec": E.shape[0] / dt,
                      "block_lanczos_steps_max": steps, "steps_mean": float(it.mean()),
                      "steps_histogram": {int(k): int(v) for k, v in zip(*np.unique(it, return_counts=True))},
                      "spmm_ms": ms, "spmm_launches": nl, "spmm_share": ms * 1e-3 / dt,
                      "launches": c1["launches"] - c0["launches"],
                      "dense_bytes_per_step_GB": step_bytes / 1e9,
                      "whole_step_GBs_dense_only": step_bytes * steps / dt / 1e9}))


if __name__ == "__main__":
    main()
