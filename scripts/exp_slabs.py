"""Column-slab experiment: is Y = A X faster as S launches Y += A_s X_s, each restricted to a
column slab of A so that the live window of the X panel is 128/S MB instead of 128 MB?  No kernel change needed:
A_s keeps only the columns of slab s (same n x n shape), so spmm(A_0) + ... + spmm(A_{S-1}) gathers exactly the
same row-tiles as spmm(A), slab by slab.  The accumulate-into-Y epilogue of the later slabs (one extra read of Y)
is NOT included: add 4.1 GB / 6.5 TB/s = 0.63 ms per extra slab at k = 512.
usage: python scripts/exp_slabs.py [S ...]"""
import json
import sys
import os

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from exp_spmm import graph, N, K  # noqa: E402


def time_spmm(ctx, lib, M, X, Y, reps=8):
    for _ in range(3):
        lib.kr_spmm_dev(ctx.h, M.h, X.h, Y.h)
    ctx.sync()
    ctx.set_timing(True)
    ctx.spmm_time(True)
    for _ in range(reps):
        lib.kr_spmm_dev(ctx.h, M.h, X.h, Y.h)
    ctx.sync()
    ms, cnt = ctx.spmm_time(True)
    ctx.set_timing(False)
    return ms / cnt


def main():
    from krylov_robustness_b200.engine import Context, Dense, Matrix
    A = graph().tocsc()
    ctx = Context.default()
    lib = ctx.lib
    X, Y = Dense(N, K, ctx).fill_rademacher(1), Dense(N, K, ctx)
    full = time_spmm(ctx, lib, Matrix(A.tocsr(), ctx), X, Y)
    print(json.dumps({"slabs": 1, "ms_total": full, "ms_each": [full]}), flush=True)
    for S in [int(a) for a in sys.argv[1:]] or [2, 4]:
        bounds = np.linspace(0, N, S + 1).astype(np.int64)
        each = []
        for s in range(S):
            lo, hi = bounds[s], bounds[s + 1]
            As = sp.csc_matrix((A.data[A.indptr[lo]:A.indptr[hi]], A.indices[A.indptr[lo]:A.indptr[hi]],
                                np.concatenate([np.zeros(lo, dtype=A.indptr.dtype), A.indptr[lo:hi + 1] - A.indptr[lo],
                                                np.full(N - hi, A.indptr[hi] - A.indptr[lo], dtype=A.indptr.dtype)])), shape=(N, N)).tocsr()
            Ms = Matrix(As, ctx)
            each.append(time_spmm(ctx, lib, Ms, X, Y))
            del Ms
        print(json.dumps({"slabs": S, "ms_total": sum(each), "ms_each": each,
                          "ms_total_plus_accumulate_estimate": sum(each) + (S - 1) * 0.63}), flush=True)


if __name__ == "__main__":
    main()
