#!/usr/bin/env python
"""How long does one greedy edge update take on the C3 graph?  kr_matrix_set_edges re-analyses and re-uploads
the matrix on the host (csr.cuh::upload_csr); one greedy round = scoring (bench secondary metric) + this."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def main():
    import krylov_robustness_b200 as kr
    from exp_spmm import graph
    A = graph()
    n = A.shape[0]
    ctx = kr.Context.default()
    t0 = time.perf_counter(); M = kr.Matrix(A, ctx); ctx.sync(); t_create = time.perf_counter() - t0
    out = {"n": n, "nnz": int(A.nnz), "matrix_create_s": t_create}
    rng = np.random.default_rng(0)
    for cnt in (1, 50):
        i = rng.integers(1, n + 1, cnt); j = rng.integers(1, n + 1, cnt)
        keep = i != j
        t0 = time.perf_counter(); M.set_edges(i[keep], j[keep], 1.0 / 64.0); ctx.sync()
        out["set_edges_%d_s" % cnt] = time.perf_counter() - t0
    print(json.dumps(out))


if __name__ == "__main__":
    main()
