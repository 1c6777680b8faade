#!/usr/bin/env python
"""Inputs of scripts/make_reference_goldens_c1*.m beyond tests/golden/reference_inputs.mat: Oregon graphs A4, A7 (the
largest) and A8 with centrality and tolerance, for BASELINE config C1 at the reference's own call shape.  Deterministic;
committed as tests/golden/reference_inputs_c1.mat.   python scripts/make_reference_inputs_c1.py"""
import os
import sys

import numpy as np
import scipy.io as sio
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_inputs():
    from conftest import load_graph
    import oracle as O
    out = {}
    for k in ("A7", "A4", "A8"):
        A = load_graph("oregon_" + k)
        out[k] = sp.csc_matrix(A)
        if k == "A7":
            out["A7_centrality"] = O.compute_centrality(A, "eig").reshape(-1, 1)
            out["A7_tol"] = 1e-6 * float(np.exp(O.normest(A, 1e-2)[0]))
    return out


def lcg_sign_probes(n, cols=680):
    """The +-1 probe block of scripts/make_reference_goldens_c1_trace.m: one Park-Miller generator per column
    (x <- 16807 x mod (2^31 - 1), seeds 1..cols after a 10-step warm-up), sign(x - 2^30) per row.  Every product stays
    below 2^53, so MATLAB doubles and NumPy integers produce the same bits."""
    x = np.arange(1, cols + 1, dtype=np.int64)
    for _ in range(10):
        x = (16807 * x) % 2147483647
    P = np.empty((n, cols))
    for i in range(n):
        x = (16807 * x) % 2147483647
        P[i] = np.where(x >= 1073741824, 1.0, -1.0)
    return P


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "reference_inputs_c1.mat")
    sio.savemat(path, build_inputs(), format="5", do_compression=True)
    print("wrote", path, os.path.getsize(path), "bytes")
