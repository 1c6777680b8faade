"""mlab - a from-scratch interpreter for the subset of MATLAB that COMPiLELab/krylov_robustness is written in.

TEST INFRASTRUCTURE ONLY (part of ``oracle/``): nothing under ``krylov_robustness_b200/`` imports it.

Why it exists.  The reference is pure MATLAB, ships no golden vectors, and this image has neither MATLAB nor GNU
Octave, so until round 2 nothing tied the NumPy/SciPy oracle (``oracle/*.py``, a hand RESTATEMENT of functions/*.m) to
outputs of the reference itself: parity was "unpinned".  This package executes the reference's OWN source files -
``/root/reference/functions/*.m``, unmodified, read where they lie - so that the control flow, the indexing, the
stopping rules, the default arguments and every quirk of the reference come from the reference's text and not from a
transcription.  ``scripts/run_reference_goldens.py`` runs ``scripts/make_reference_goldens.m`` (the very script a
maintainer with Octave / MATLAB would run) through it and commits the result as
``tests/golden/reference_golden.json``; ``tests/test_reference_goldens.py`` then checks the oracle (CPU tier) and
the device path (GPU tier) against that file at 1e-10 with equal iteration counts.

What it is NOT.  MATLAB's built-ins are closed source.  Here ``qr``, ``eig``, ``expm``, sparse products ... are NumPy /
SciPy calls into the same LAPACK routines (``oracle/mlab/builtins.py``), and ``normest`` / ``normest1`` (MathWorks
m-files that are not part of the reference repository) are restated from their published algorithms.  Rounding can
therefore differ from a real MATLAB session at the 1e-15 level; iteration counts, flags and selected edges are what
such a session would produce unless a stopping test falls within rounding of its threshold.

Language coverage (what functions/*.m and the generator script need): functions with sub-functions and nested
functions (shared workspace), varargin / nargin / nargout, multiple return values with ``~`` placeholders, value
semantics, anonymous functions capturing their workspace, function handles, cell arrays with comma-separated-list
expansion (``c{:}``), scalar structs with dynamic fields, ``end`` arithmetic inside indices, growth on indexed
assignment, nested lvalues (``X{h}{1}(n, n) = 0``), logical / linear / 2-D indexing on full and sparse matrices,
implicit expansion, if / for / while / switch, command syntax (``load theta_taylor``), globals, file I/O for the
JSON writer.
"""
from .interp import Interpreter
from .values import MatlabError, Cell, Struct, FH

__all__ = ["Interpreter", "MatlabError", "Cell", "Struct", "FH"]
