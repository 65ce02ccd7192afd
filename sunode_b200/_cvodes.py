"""Compatibility shim for the raw SUNDIALS pokes shown in the reference's README (:240-250):

    lib = sunode._cvodes.lib
    lib.CVodeSStolerancesB(solver._ode, solver._odeB, 1e-8, 1e-8)
    lib.CVodeQuadSStolerancesB(solver._ode, solver._odeB, 1e-8, 1e-8)
    lib.CVodeSetMaxNumSteps(solver._ode, 5000)
    lib.CVodeSetMaxNumStepsB(solver._ode, solver._odeB, 5000)

There is no CVODES memory block behind ``solver._ode`` / ``solver._odeB`` here; they are tokens
that carry the solver, and the five functions forward to its setters.
"""
from __future__ import annotations

import numpy as np


class OdeToken:
    """What ``solver._ode`` / ``solver._odeB`` evaluate to."""
    __slots__ = ('solver', 'backward')

    def __init__(self, solver, backward: bool):
        self.solver = solver
        self.backward = backward


class _Lib:
    CV_BDF = 2
    CV_ADAMS = 1
    CV_SUCCESS = 0
    CV_TOO_MUCH_WORK = -1

    @staticmethod
    def CVodeSStolerances(ode: OdeToken, reltol: float, abstol: float) -> int:
        ode.solver._set_tolerances(np.float64(abstol), np.float64(reltol))
        return 0

    @staticmethod
    def CVodeSStolerancesB(ode: OdeToken, odeB: OdeToken, reltol: float, abstol: float) -> int:
        odeB.solver.set_backward_tolerances(reltol, abstol)
        return 0

    @staticmethod
    def CVodeQuadSStolerancesB(ode: OdeToken, odeB: OdeToken, reltol: float, abstol: float) -> int:
        odeB.solver.set_quad_tolerances(reltol, abstol)
        return 0

    @staticmethod
    def CVodeSetMaxNumSteps(ode: OdeToken, mxsteps: int) -> int:
        ode.solver.set_max_num_steps(int(mxsteps))
        return 0

    @staticmethod
    def CVodeSetMaxNumStepsB(ode: OdeToken, odeB: OdeToken, mxsteps: int) -> int:
        odeB.solver.set_max_num_steps_backward(int(mxsteps))
        return 0


lib = _Lib()
