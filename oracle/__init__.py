"""CPU oracle for the Krylov matrix-function hot path of COMPiLELab/krylov_robustness.

TEST INFRASTRUCTURE ONLY.  Nothing under ``krylov_robustness_b200/`` imports this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may.  It is the checker, never the product.

What it is: a NumPy/SciPy restatement, function by function, of the reference's MATLAB files
under ``functions/`` (each function's docstring cites the file:line it follows).  The reference
is pure MATLAB and neither MATLAB nor GNU Octave exists in this image, so the reference itself
cannot be executed here.

PARITY PINNED (round 2): ``oracle/mlab`` is a from-scratch interpreter for the MATLAB subset the reference is
written in; ``scripts/run_reference_goldens.py`` executes the reference's UNMODIFIED ``functions/*.m`` with it (driven
by ``scripts/make_reference_goldens.m``) and the outputs are committed as ``tests/golden/reference_golden.json`` with
their provenance (SHA-256 of every reference file executed).  ``tests/test_reference_goldens.py`` holds this package
and the device path to those outputs: 1e-10 relative, iteration counts / flags / selected edges equal.  MATLAB's
closed-source built-ins are LAPACK / SciPy stand-ins there as here (``oracle/mlab/__init__.py`` says what that means).
The dense identities the reference keeps in its ``debug`` branches (``trace_fun_update.m:91-102``,
``fun_and_grad_krylov_exp.m:90-111``, ``function_multiple_entries.m:80-82``) remain as a second, independent pin:
tests/test_oracle_*.py turn them into assertions against dense ``eigh``/``expm`` ground truth on the reference's graphs.
"""
from .theta import THETA
from .krylov import lanczos_krylov, arnoldi_krylov
from .updates import (trace_fun_update, fun_update, function_multiple_entries,
                      fun_and_grad_krylov_exp, fun_and_grad_krylov_fun, normest)
from .expmv import expmv, select_taylor_degree, normAm
from .mctrace import mc_trace, trace_exp, slq_trace
from .frechet import multiple_frechet_eval, hessianfcn_exp, hessianfcn_fun
from .greedy import (krylov_miobi, greedy_krylov, find_top_edges, find_top_missing_edges,
                     edge2low_rank, compute_centrality)

__all__ = [
    "THETA", "lanczos_krylov", "arnoldi_krylov", "trace_fun_update", "fun_update",
    "function_multiple_entries", "fun_and_grad_krylov_exp", "fun_and_grad_krylov_fun",
    "normest", "expmv", "select_taylor_degree", "normAm", "mc_trace", "trace_exp",
    "slq_trace", "krylov_miobi", "greedy_krylov", "find_top_edges",
    "find_top_missing_edges", "edge2low_rank", "compute_centrality",
    "multiple_frechet_eval", "hessianfcn_exp", "hessianfcn_fun",
]
