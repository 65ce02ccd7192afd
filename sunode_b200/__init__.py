"""sunode_b200: batched stiff-ODE + adjoint engine for B200 behind sunode's API.

Import surface of the reference kept (``sunode/__init__.py``): ``SympyProblem``, ``solver``
(``Solver``, ``AdjointSolver``, ``SolverError``), ``wrappers.as_pytensor``, ``_cvodes.lib``.
"""
from . import _cvodes, basic, dtypesubset, problem, symode
from . import solver, wrappers  # noqa: E402  (the C-ABI library itself loads lazily, at first use)
from .basic import SolverError
from .symode import SympyProblem

__version__ = "0.1.0"

__all__ = ["SympyProblem", "SolverError", "symode", "basic", "dtypesubset", "problem", "_cvodes",
           "solver", "wrappers"]
