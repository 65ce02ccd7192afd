from .problem import SympyProblem

__all__ = ['SympyProblem']
