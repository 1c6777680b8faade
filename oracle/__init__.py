"""CPU oracle for the Krylov matrix-function hot path of COMPiLELab/krylov_robustness.

TEST INFRASTRUCTURE ONLY.  Nothing under ``krylov_robustness_b200/`` imports this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may.  It is the checker, never the product.

What it is: a NumPy/SciPy restatement, function by function, of the reference's MATLAB files
under ``functions/`` (each function's docstring cites the file:line it follows).  The reference
is pure MATLAB and neither MATLAB nor GNU Octave exists in this image, so the reference itself
cannot be executed here.

PARITY UNPINNED: the reference ships no golden vectors, no assertions and no expected-value
files.  The only pins available are the dense identities the reference
keeps in its ``debug`` branches (``trace_fun_update.m:91-102``, ``fun_and_grad_krylov_exp.m:90-111``,
``function_multiple_entries.m:80-82``) - tests/test_oracle_*.py turn those into assertions against
dense ``eigh``/``expm`` ground truth on the reference's own graphs.  MATLAB built-ins with no
source (qr, eig, expm, funm, normest, eigs) are replaced by their LAPACK / SciPy equivalents.
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
