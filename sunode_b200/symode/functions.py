"""Extra sympy functions understood by the code generator.

Counterparts of the helper functions the reference makes available inside generated code
(``sunode/symode/lambdify.py:59-77`` for the numeric definitions, ``:275-352`` for the sympy
classes): ``logaddexp``, ``expit``, ``dexpit``, ``CardinalBSpline(4, t)`` and
``interpolate_spline``.  Each has a device/C implementation in :mod:`.codegen`'s prelude.

The reference's ``expit.fdiff``/``dexpit.fdiff`` test ``argindex`` the wrong way round and
therefore raise whenever sympy differentiates through them (lambdify.py:301-305,318-322); that
is an upstream bug, not behaviour, so the derivatives are implemented correctly here.
"""
from __future__ import annotations

from functools import partial

import sympy as sy
from sympy.core.function import ArgumentIndexError


class logaddexp(sy.Function):
    """log(exp(a) + exp(b)), evaluated stably in generated code."""
    nargs = 2

    def fdiff(self, argindex=1):
        if argindex not in (1, 2):
            raise ArgumentIndexError(self, argindex)
        a, b = self.args
        return sy.exp(self.args[argindex - 1]) / (sy.exp(a) + sy.exp(b))

    def _eval_is_real(self):
        return self.args[0].is_real and self.args[1].is_real

    def _eval_is_finite(self):
        return self.args[0].is_finite and self.args[1].is_finite


class expit(sy.Function):
    """Logistic function 1 / (1 + exp(-x))."""
    nargs = 1

    def fdiff(self, argindex=1):
        if argindex != 1:
            raise ArgumentIndexError(self, argindex)
        return dexpit(self.args[0])

    def _eval_is_real(self):
        return self.args[0].is_real


class dexpit(sy.Function):
    """Derivative of the logistic function, expit(x) * expit(-x)."""
    nargs = 1

    def fdiff(self, argindex=1):
        if argindex != 1:
            raise ArgumentIndexError(self, argindex)
        x = self.args[0]
        return dexpit(x) * (1 - 2 * expit(x))

    def _eval_is_real(self):
        return self.args[0].is_real


class CardinalBSpline(sy.Function):
    """Cardinal B-spline basis function of the given degree on knots 0..degree+1.

    Only degree 4 has a generated-code implementation (as in the reference,
    lambdify.py:73-77, which returns NaN for other degrees)."""
    nargs = 2

    def as_sympy_expr(self):
        degree, x = self.args
        knots = tuple(sy.Integer(i) for i in range(int(degree) + 2))
        basis = sy.functions.special.bsplines.bspline_basis(int(degree), knots, 0, x)
        return sy.Piecewise(*[(sy.horner(expr), cond) for expr, cond in basis.args])


def interpolate_spline(x, vals, lower, upper, degree, as_pure=False):
    """Spline with coefficients ``vals`` on [lower, upper] (reference lambdify.py:343-352)."""
    n_vals = len(vals)
    n_knots = degree + n_vals + 1
    basis = partial(CardinalBSpline, degree)
    x = (x - lower) / (upper - lower)
    x = degree + x * (n_knots - 2 * degree - 1)
    terms = [basis(x - i) for i in range(n_vals)]
    if as_pure:
        terms = [b.as_sympy_expr() for b in terms]
    return sum(val * b for val, b in zip(vals, terms))
