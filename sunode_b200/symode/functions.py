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


# ---------------------------------------------------------------------------------------------
# Rewrites for numerically stable soft-max style expressions -- counterparts of the reference's
# ``logsumexp_2terms_opt`` / ``simplify_multiple_exp_sum`` / ``explog_opt``
# (sunode/symode/lambdify.py:355-432).  They are sympy ``ReplaceOptim`` objects, meant to be
# applied through ``SympyProblem(..., simplify=lambda e: optimize(e, [explog_opt]))``:
#
#   log(exp(a) + exp(b))                     ->  logaddexp(a, b)
#   exp(b) / (exp(a) + exp(b)) ** n * ...    ->  +-exp(b - n * logaddexp(a, b) + ...)
#
# i.e. a product with more than one factor built from exp() sums, whose sign is known, is moved
# into log space, where the sums become ``logaddexp`` and cannot overflow.
from sympy.assumptions import Q, ask                      # noqa: E402
from sympy.codegen.rewriting import ReplaceOptim, log1p_opt, optimize   # noqa: E402


def _two_exps(expr) -> bool:
    return isinstance(expr, sy.Add) and len(expr.args) == 2 and all(isinstance(a, sy.exp) for a in expr.args)


def _log_of_two_exps(expr) -> bool:
    return isinstance(expr, sy.log) and _two_exps(expr.args[0])


logsumexp_2terms_opt = ReplaceOptim(
    _log_of_two_exps,
    lambda e: logaddexp(e.args[0].args[0].args[0], e.args[0].args[1].args[0]))


def is_exp_sum(expr) -> bool:
    """``exp(a)`` or ``exp(a) + exp(b)``."""
    return isinstance(expr, sy.exp) or _two_exps(expr)


def is_exp_sum_pow(expr) -> bool:
    return is_exp_sum(expr) or (isinstance(expr, sy.Pow) and is_exp_sum(expr.args[0]))


def is_exp_sum_pow_mult(expr) -> bool:
    return is_exp_sum_pow(expr) or (isinstance(expr, sy.Mul) and any(is_exp_sum_pow(a) for a in expr.args))


def is_multiple_exp_sum_pow_mult(expr) -> bool:
    return isinstance(expr, sy.Mul) and sum(1 for a in expr.args if is_exp_sum_pow_mult(a)) > 1


def _sign_of(expr):
    if ask(Q.positive(expr)):
        return 1
    if ask(Q.negative(expr)):
        return -1
    return None


def simplify_multiple_exp_sum(expr, do_simplify=False, optims=None):
    """Move a product of known sign into log space: ``s * exp(expand_log(log(s * expr)))`` with
    the logarithms of exp() sums rewritten as ``logaddexp`` / ``log1p``; expressions of unknown
    sign are searched for such products argument by argument."""
    if optims is None:
        optims = (log1p_opt, logsumexp_2terms_opt)
    sign = _sign_of(expr)
    if sign is None:
        if not expr.args:
            return expr
        return expr.func(*[simplify_multiple_exp_sum(a, do_simplify, optims) for a in expr.args])
    # (expand_log does not see assumptions made through a context manager: force)
    in_log_space = optimize(sy.expand_log(sy.log(sign * expr), force=True), optims)
    return sign * sy.exp(in_log_space, evaluate=False)


explog_opt = ReplaceOptim(
    lambda e: _sign_of(e) is not None and is_multiple_exp_sum_pow_mult(e),
    simplify_multiple_exp_sum)
