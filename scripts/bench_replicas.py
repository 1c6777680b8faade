#!/usr/bin/env python
"""One process, all visible GPUs, nothing but the C ABI (kr_matrix_replicate): candidate-edge scoring (config C5
shape on the C3 graph) and the host-probe SLQ pass on 1 GPU and on the replicated matrix."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import krylov_robustness_b200 as kr
    from krylov_robustness_b200.graphs import power_law_graph, spectral_radius_estimate
    n = int(os.environ.get("KR_BENCH_N", 1_000_000))
    nnz = int(os.environ.get("KR_BENCH_NNZ", 20_000_000))
    ncand = int(os.environ.get("KR_BENCH_C5_CAND", 16384))
    k = int(os.environ.get("KR_BENCH_K", 512))
    A = power_law_graph(n, nnz, 2.2, 20260310)
    lam = spectral_radius_estimate(A, 30)
    A = (A * (1.0 / lam)).tocsr()
    M = kr.Matrix(A)
    cvec = kr.compute_centrality(M, "eig", 1e-10)
    E = kr.find_top_missing_edges(A, cvec, ncand, "min")
    tol = 1e-6 * float(np.e)
    Z = torch.empty((k, n), dtype=torch.int8, pin_memory=True)
    Z.numpy()[:] = kr.rademacher_host(n, k, 1).T.astype(np.int8)
    Zc = Z.numpy().T                                       # column-major n x k view of the pinned buffer

    def run():
        kr.trace_fun_update_edges(M, E[:256], 1.0 / lam, tol, 100, "exp")
        t0 = time.perf_counter()
        x, it, _ = kr.trace_fun_update_edges(M, E, 1.0 / lam, tol, 100, "exp")
        t_edges = time.perf_counter() - t0
        kr.slq_trace(M, Zc, 4, "exp")
        t0 = time.perf_counter()
        tr = kr.slq_trace(M, Zc, 30, "exp")
        t_slq = time.perf_counter() - t0
        return x, it, t_edges, tr, t_slq
    x1, it1, te1, tr1, ts1 = run()
    R = M.replicate()
    xr, itr, ter, trr, tsr = run()
    print(json.dumps({"gpus_visible": torch.cuda.device_count(), "replicas": R, "candidates": int(E.shape[0]),
                      "edges_per_s_1gpu": E.shape[0] / te1, "edges_per_s_replicated": E.shape[0] / ter,
                      "edges_speedup": te1 / ter, "scores_identical": bool(np.array_equal(x1, xr) and np.array_equal(it1, itr)),
                      "slq_e2e_matvecs_per_s_1gpu": k * 30 / ts1, "slq_e2e_matvecs_per_s_replicated": k * 30 / tsr,
                      "slq_speedup": ts1 / tsr, "trace_rel_diff": abs(trr - tr1) / abs(tr1),
                      "note": "one process, host threads behind the C ABI (no torch.distributed, no NCCL): the exchange step is "
                              "the host-side gather of scores / sum of partial traces"}))


if __name__ == "__main__":
    main()
