"""CPU-side checks of the drop-in boundary: the shared library loads, exports every function the
header declares (no compute calls without a GPU), the ctypes table matches the header, and the product
path refuses to run without a device instead of falling back to anything on the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "krylov_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kr_[a-zA-Z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from krylov_robustness_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 35
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.PROTOTYPES) == names          # the binding table and the header agree


def test_every_entry_point_cites_the_reference():
    src = open(HEADER).read()
    for ref in ("functions/lanczos_krylov.m", "functions/arnoldi_krylov.m", "functions/trace_fun_update.m",
                "functions/fun_update.m", "functions/function_multiple_entries.m",
                "functions/fun_and_grad_krylov_exp.m", "functions/fun_and_grad_krylov_fun.m",
                "functions/mc_trace.m", "functions/expmv.m", "functions/select_taylor_degree.m",
                "functions/normAm.m", "functions/krylov_miobi.m", "functions/theta_taylor.mat"):
        assert ref in src, ref


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import krylov_robustness_b200 as kr
    with pytest.raises(kr._lib.KrylovB200Error, match="no CPU fallback"):
        kr.Context(0)
    import scipy.sparse as sp
    with pytest.raises(kr._lib.KrylovB200Error):
        kr.trace_fun_update(sp.identity(200, format="csr"), np.ones((200, 1)), [[1.0]])


def test_theta_table_matches_oracle_without_gpu():
    from krylov_robustness_b200 import _lib
    import oracle
    th = (ctypes.c_double * 100)()
    assert _lib.load().kr_theta(th) == 0
    assert np.array_equal(np.array(th), oracle.THETA)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "krylov_robustness_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".h", ".c", ".m")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text, f
