"""Loading the reference's data files the way its experiment scripts do (row f4 of SURVEY.md section 8).

Host-side glue only (SciPy): MAT-v5 files are read with scipy.io.loadmat; the three MAT-v7.3 (HDF5) files of
datasets_paper/Misc (CollegeMsg, Drugs, as_735) go through the minimal HDF5 reader of hdf5_min.py (the image has no
HDF5 library).  Preprocessing mirrors the scripts line by line:

* unweighted experiments (Tests/test_unweighted_break.m:45-53,160-169): A <- spones(A + A'), zero diagonal,
  largest connected component (first largest label);
* weighted experiments (Tests/test_weighted_sinh_lbfgs.m:48-52): one symmetric weighted adjacency per country
  in voltage_adjacencies_average_2.mat, A <- A / max(A).
"""
import os

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components


def _loadmat(path):
    import scipy.io as sio
    from . import hdf5_min
    if hdf5_min.is_v73(path):
        return hdf5_min.loadmat73(path)
    try:
        return sio.loadmat(path, spmatrix=True)
    except TypeError:                      # older SciPy: no spmatrix keyword
        return sio.loadmat(path)
    except NotImplementedError as e:       # MAT v7.3
        raise ValueError("%s is a MAT-v7.3 (HDF5) file; no HDF5 reader is available here" % os.path.basename(path)) from e


def largest_component(A):
    """Sub-matrix on the largest connected component; ties go to the first label, as the reference's loop does
    (Tests/test_unweighted_break.m:160-169)."""
    _, lab = connected_components(A, directed=False)
    big = int(np.argmax(np.bincount(lab)))
    idx = np.where(lab == big)[0]
    return sp.csr_matrix(A)[idx][:, idx].tocsr()


def unweighted_adjacency(A):
    """spones(A + A') without self loops, restricted to the largest component
    (Tests/test_unweighted_break.m:45-53)."""
    A = sp.csr_matrix(A).astype(np.float64)
    A = (A + A.T).tocsr()
    A.data[:] = 1.0
    A.setdiag(0)
    A.eliminate_zeros()
    A = largest_component(A)
    A.sort_indices()
    return A


def load_problem(path, unweighted=True):
    """A SuiteSparse-style .mat (`Problem.A`, datasets_paper/Misc and datasets_paper/Transport) or a file holding
    plain sparse variables (MIOBI Codes/dt_oregon.mat: pass `path::A0`)."""
    var = None
    if "::" in path:
        path, var = path.split("::", 1)
    m = _loadmat(path)
    if var is not None:
        A = m[var]
    elif "Problem" in m:
        A = m["Problem"]["A"] if isinstance(m["Problem"], dict) else m["Problem"]["A"][0, 0]
    else:
        names = [k for k in m if not k.startswith("__")]
        if len(names) != 1:
            raise ValueError("%s holds %s; pick one with path::name" % (os.path.basename(path), names))
        A = m[names[0]]
    A = sp.csr_matrix(A).astype(np.float64)
    return unweighted_adjacency(A) if unweighted else A


def load_power_grids(path, countries=None):
    """{country: A / max(A)} from voltage_adjacencies_average_2.mat (Tests/test_weighted_sinh_lbfgs.m:40-52)."""
    m = _loadmat(path)
    out = {}
    for k in (countries or [k for k in m if not k.startswith("__")]):
        A = sp.csr_matrix(m[k]).astype(np.float64)
        if (A != A.T).nnz:
            A = (A + A.T) / 2
        A = (A / A.max()).tocsr()
        A.sort_indices()
        A.eliminate_zeros()
        out[k] = A
    return out
