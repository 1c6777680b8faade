#!/usr/bin/env python
"""Inputs of scripts/make_reference_goldens_c1.m beyond tests/golden/reference_inputs.mat: the largest Oregon graph
(A7) with its centrality and tolerance, for BASELINE config C1 at the reference's own call shape.  Deterministic;
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
    A7 = load_graph("oregon_A7")
    c7 = O.compute_centrality(A7, "eig")
    return {"A7": sp.csc_matrix(A7), "A7_centrality": c7.reshape(-1, 1),
            "A7_tol": 1e-6 * float(np.exp(O.normest(A7, 1e-2)[0]))}


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "reference_inputs_c1.mat")
    sio.savemat(path, build_inputs(), format="5", do_compression=True)
    print("wrote", path, os.path.getsize(path), "bytes")
