"""sunode_b200: batched stiff-ODE + adjoint engine for B200 behind sunode's API."""
from .symode import SympyProblem
from .basic import SolverError

__version__ = "0.1.0"

__all__ = ["SympyProblem", "SolverError"]
