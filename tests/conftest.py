import os
import sys
import warnings

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    # The library is a build artefact (git-ignored).  On a fresh checkout with a CUDA toolkit, compile it once
    # (nvcc cross-compiles sm_100a without a GPU, ~1.5 min) so that the C-ABI / MEX-gateway tests have
    # something to load; on the GPU box the prebuilt .so travels with the snapshot.
    lib = os.path.join(ROOT, "krylov_robustness_b200", "libkrylov_b200.so")
    if not os.path.exists(lib) and os.path.exists("/usr/local/cuda/bin/nvcc"):
        import subprocess
        subprocess.run(["bash", os.path.join(ROOT, "krylov_robustness_b200", "csrc", "build.sh")], check=False)
    warnings.filterwarnings("ignore", message=".*Reached maximum number of iterations.*")
    warnings.filterwarnings("ignore", message=".*lucky breakdown.*")


def load_graph(name):
    z = np.load(os.path.join(GOLDEN, "graph_%s.npz" % name))
    n = int(z["n"])
    return sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))


@pytest.fixture(scope="session")
def graphs():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_graph(name)
        return cache[name]
    return get


def edge_UB(n, i, j, sign):
    """U, B of functions/krylov_miobi.m:77-98 for one candidate edge (1-based i, j)."""
    if i != j:
        U = np.zeros((n, 2))
        U[i - 1, 0] = 1.0
        U[j - 1, 1] = 1.0
        return U, sign * np.array([[0.0, 1.0], [1.0, 0.0]])
    U = np.zeros((n, 1))
    U[i - 1, 0] = 1.0
    return U, np.array([[sign]])
