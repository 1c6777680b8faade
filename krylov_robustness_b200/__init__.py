"""krylov_robustness_b200 - B200-native engine behind the function interface of
COMPiLELab/krylov_robustness (Krylov matrix-function hot path only).

Importing the package loads libkrylov_b200.so; there is no CPU fallback.
"""
from . import _lib
from .engine import Context, Matrix, Dense, rademacher_host
from .functions import *  # noqa: F401,F403

__all__ = ["Context", "Matrix", "Dense", "rademacher_host"]
