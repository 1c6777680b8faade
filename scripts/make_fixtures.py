"""Build the small graph fixtures under tests/golden/ from the reference's data files.

Run in the build container only (needs /root/reference):  python scripts/make_fixtures.py
The preprocessing mirrors Tests/test_unweighted_break.m:45-53 (symmetrise with spones(A+A'),
drop self-loops, largest connected component) and Tests/test_weighted_sinh_lbfgs.m:48 (A/max(A)).
Each fixture is a CSR triple (indptr, indices, data) + n in one .npz (a few kB each).
"""
import os

import sys

import numpy as np
import scipy.io as sio
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components

REF = "/root/reference"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def save(name, A):
    A = sp.csr_matrix(A).astype(np.float64)
    A.sort_indices()
    A.eliminate_zeros()
    assert (A != A.T).nnz == 0
    idt = np.int32 if A.shape[0] > 50000 else np.int64          # loaders cast back; keeps the big fixture small
    np.savez_compressed(os.path.join(OUT, "graph_%s.npz" % name), n=A.shape[0], indptr=A.indptr.astype(idt),
                        indices=A.indices.astype(idt), data=A.data)
    print(name, A.shape[0], A.nnz)


from krylov_robustness_b200.datasets import largest_component as lcc          # noqa: E402,F401
from krylov_robustness_b200.datasets import unweighted_adjacency as unweighted   # noqa: E402


def main():
    os.makedirs(OUT, exist_ok=True)
    m = sio.loadmat(os.path.join(REF, "MIOBI Codes", "dt_oregon.mat"), spmatrix=True)
    for k in ("A0", "A1", "A2", "A3", "A4", "A5", "A6", "A7", "A8"):       # all nine (config C1 names dt_oregon.mat as a whole)
        save("oregon_%s" % k, m[k])
    for k in ("Anaheim", "Barcelona", "Rome", "Vermont"):       # Vermont = the largest road network (config C2 at full size)
        p = sio.loadmat(os.path.join(REF, "datasets_paper", "Transport", k + ".mat"), spmatrix=True)
        save("transport_%s" % k, unweighted(p["Problem"]["A"][0, 0]))
    v = sio.loadmat(os.path.join(REF, "datasets_paper", "voltage_adjacencies_average_2.mat"), spmatrix=True)
    for k in ("Austria", "Sweden", "England", "Mexico"):
        A = sp.csr_matrix(v[k]).astype(np.float64)
        A = (A + A.T) / 2 if (A != A.T).nnz else A
        save("grid_%s" % k, A / A.max())
    # the three MAT-v7.3 (HDF5) files of datasets_paper/Misc, through the minimal HDF5 reader (hdf5_min.py)
    from krylov_robustness_b200.datasets import load_problem
    for k in ("Drugs", "CollegeMsg", "as_735"):
        save("misc_%s" % k, load_problem(os.path.join(REF, "datasets_paper", "Misc", k + ".mat")))


if __name__ == "__main__":
    main()
