"""sympy -> CUDA ``__device__`` / C99 source for the problem functions.

This replaces the back half of the reference's ``sunode/symode/lambdify.py`` (``lambdify_consts``
/ ``LambdifyAST``, lines 38-186 and 203-270), which turns the same sympy arrays into Python ASTs
for numba.  The contract kept from the reference:

* common sub-expression elimination per function (lambdify.py:253-256);
* structurally-zero entries are written as zeros and cost nothing (lambdify.py:102-127);
* dense Jacobians are stored **column-major** (problem.py:345,377 use ``numba.farray``);
* the helper functions ``logaddexp / expit / dexpit / CardinalBSpline(4, .)`` exist
  (lambdify.py:59-77).

What differs by design: the target is a flat-array calling convention

    sb_rhs     (t, y[NS], p[NP], out[NS])
    sb_jac     (t, y, p, J[NS*NS])            J[i + NS*j] = d f_i / d y_j
    sb_adj_rhs (t, y, lam[NS], p, out[NS])    -J^T lam        (symode/problem.py:147)
    sb_adj_jac (t, y, p, JB[NS*NS])           -J^T            (symode/problem.py:410)
    sb_quad_rhs(t, y, lam, p, out[ND])        lam^T df/dp     (symode/problem.py:148)
    sb_sens_rhs(t, y, s[ND*NS], p, out[ND*NS]) J s_k + df/dp_k (symode/problem.py:512-513)

emitted twice from one expression set: as ``__device__ __forceinline__`` functions that are
compiled *into* the sm_100a integrator kernels (every index is a literal, so all arrays stay in
registers), and as plain C with exported wrappers for host-side evaluation.  Non-finite detection
(the reference's ``return 1``) is done by the callers, not inside the generated bodies.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np
import sympy as sy
from sympy.printing.c import C99CodePrinter

CODEGEN_VERSION = 3


class _Printer(C99CodePrinter):
    """C99 printer with cheap integer powers and the helper functions mapped to ``sb_*``."""

    def __init__(self) -> None:
        super().__init__({
            'user_functions': {
                'logaddexp': 'sb_logaddexp',
                'expit': 'sb_expit',
                'dexpit': 'sb_dexpit',
                'CardinalBSpline': 'sb_cardinal_bspline',
            },
            'allow_unknown_functions': False,
        })

    def _print_Pow(self, expr):  # type: ignore[override]
        base, exp = expr.args
        if exp.is_Integer:
            n = int(exp)
            if 1 <= abs(n) <= 4:
                b = self.parenthesize(base, sy.printing.precedence.PRECEDENCE['Mul'] + 1)
                prod = '*'.join([b] * abs(n))
                if n > 0:
                    return '(%s)' % prod
                return '(1.0/(%s))' % prod
        if exp == sy.Rational(1, 2):
            return 'sqrt(%s)' % self._print(base)
        if exp == -sy.Rational(1, 2):
            return '(1.0/sqrt(%s))' % self._print(base)
        return super()._print_Pow(expr)

    def _print_Integer(self, expr):  # type: ignore[override]
        # integers appearing as operands of floating expressions: keep them as doubles so that
        # e.g. 2*y is a double multiply and 1/3 never becomes integer division
        return '%d.0' % int(expr) if abs(int(expr)) < 2 ** 53 else repr(float(expr))

    def _print_Rational(self, expr):  # type: ignore[override]
        return '(%d.0/%d.0)' % (expr.p, expr.q)


_HELPERS = r"""
SB_FN double sb_logaddexp(double a, double b) {
    const double lo = fmin(a, b), hi = fmax(a, b);
    return hi + log1p(exp(lo - hi));
}
SB_FN double sb_expit(double x) { return 1.0 / (1.0 + exp(-x)); }
SB_FN double sb_dexpit(double x) { return sb_expit(x) * sb_expit(-x); }
SB_FN double sb_cardinal_bspline(double degree, double t) {
    if (degree != 4.0) return NAN;
    if (t >= 0.0 && t <= 1.0) return (1.0/24.0)*t*t*t*t;
    if (t >= 1.0 && t <= 2.0) return t*(t*(t*(5.0/6.0 - 1.0/6.0*t) - 5.0/4.0) + 5.0/6.0) - 5.0/24.0;
    if (t >= 2.0 && t <= 3.0) return t*(t*(t*((1.0/4.0)*t - 5.0/2.0) + 35.0/4.0) - 25.0/2.0) + 155.0/24.0;
    if (t >= 3.0 && t <= 4.0) return t*(t*(t*(5.0/2.0 - 1.0/6.0*t) - 55.0/4.0) + 65.0/2.0) - 655.0/24.0;
    if (t >= 4.0 && t <= 5.0) return t*(t*(t*((1.0/24.0)*t - 5.0/6.0) + 25.0/4.0) - 125.0/6.0) + 625.0/24.0;
    return 0.0;
}
"""


@dataclass
class GeneratedSource:
    """Generated source for one problem plus the sizes the integrator is specialised on."""
    n_states: int
    n_params: int          # all parameters (flat), NOT the reference's n_params (= n_deriv)
    n_deriv: int
    body: str              # target independent function bodies (uses SB_FN)
    deriv_index: Tuple[int, ...] = ()
    extra: Dict[str, str] = field(default_factory=dict)

    @property
    def digest(self) -> str:
        h = hashlib.sha256()
        h.update(('v%d|%d|%d|%d|' % (CODEGEN_VERSION, self.n_states, self.n_params,
                                     self.n_deriv)).encode())
        h.update(self.body.encode())
        return h.hexdigest()[:24]

    def _sizes(self) -> str:
        return ('#define SB_NS %d\n#define SB_NP %d\n#define SB_ND %d\n'
                % (self.n_states, self.n_params, self.n_deriv))

    @property
    def cuda(self) -> str:
        """Device functions; the integrator kernels are appended by the C-ABI library."""
        return (self._sizes()
                + '#define SB_FN static __device__ __forceinline__\n'
                + self.body)

    @property
    def c(self) -> str:
        """Plain C with exported wrappers (``sbh_*``) for host evaluation / CPU callbacks."""
        ns, nd = self.n_states, self.n_deriv
        wrappers = r"""
static int sb_allfinite(const double* v, int n) {
    for (int i = 0; i < n; ++i) if (!isfinite(v[i])) return 0;
    return 1;
}
int sbh_sizes(int* ns, int* np_, int* nd) { *ns = SB_NS; *np_ = SB_NP; *nd = SB_ND; return 0; }
int sbh_rhs(double t, const double* y, const double* p, double* out) {
    sb_rhs(t, y, p, out); return !sb_allfinite(out, SB_NS);
}
int sbh_jac(double t, const double* y, const double* p, double* out) {
    sb_jac(t, y, p, out); return !sb_allfinite(out, SB_NS * SB_NS);
}
int sbh_adj_rhs(double t, const double* y, const double* lam, const double* p, double* out) {
    sb_adj_rhs(t, y, lam, p, out); return !sb_allfinite(out, SB_NS);
}
int sbh_adj_jac(double t, const double* y, const double* p, double* out) {
    sb_adj_jac(t, y, p, out); return !sb_allfinite(out, SB_NS * SB_NS);
}
int sbh_quad_rhs(double t, const double* y, const double* lam, const double* p, double* out) {
    sb_quad_rhs(t, y, lam, p, out); return !sb_allfinite(out, SB_ND);
}
int sbh_sens_rhs(double t, const double* y, const double* s, const double* p, double* out) {
    sb_sens_rhs(t, y, s, p, out); return !sb_allfinite(out, SB_ND * SB_NS);
}
"""
        return ('#include <math.h>\n' + self._sizes()
                + '#define SB_FN static inline\n' + self.body + wrappers)


def _emit_function(
    printer: _Printer,
    name: str,
    args: Sequence[str],
    out_name: str,
    exprs: Sequence[sy.Expr],
    out_index: Sequence[int],
    n_out: int,
    subs: Dict[sy.Symbol, sy.Symbol],
    tag: str,
) -> str:
    """One ``SB_FN void name(args..., double* out)`` with CSE'd body."""
    exprs = [sy.sympify(e).xreplace(subs) for e in exprs]
    tmp_names = sy.numbered_symbols('sb%s' % tag)
    replacements, reduced = sy.cse(exprs, symbols=tmp_names, order='none')
    lines: List[str] = []
    for sym, sub_expr in replacements:
        lines.append('    const double %s = %s;' % (sym.name, printer.doprint(sub_expr)))
    written = set()
    for idx, expr in zip(out_index, reduced):
        written.add(idx)
        if expr == 0:
            lines.append('    %s[%d] = 0.0;' % (out_name, idx))
        else:
            lines.append('    %s[%d] = %s;' % (out_name, idx, printer.doprint(expr)))
    for idx in range(n_out):
        if idx not in written:
            lines.append('    %s[%d] = 0.0;' % (out_name, idx))
    if not lines:
        lines.append('    (void)%s;' % out_name)
    sig = ', '.join(list(args) + ['double* __restrict__ %s' % out_name])
    unused = ''.join('    (void)%s;\n' % a.split()[-1].strip('*') for a in args)
    return 'SB_FN void %s(%s) {\n%s%s\n}\n' % (name, sig, unused, '\n'.join(lines))


def _uses_helpers(exprs: Iterable[sy.Expr]) -> bool:
    names = {'logaddexp', 'expit', 'dexpit', 'CardinalBSpline'}
    for e in exprs:
        for f in sy.sympify(e).atoms(sy.Function):
            if type(f).__name__ in names:
                return True
    return False


def generate(
    *,
    time: sy.Symbol,
    states: Sequence[sy.Symbol],
    params: Sequence[sy.Symbol],
    lamda: Sequence[sy.Symbol],
    sens: np.ndarray,
    deriv_index: Sequence[int],
    dydt: Sequence[sy.Expr],
    jac: np.ndarray,
    dydp: np.ndarray,
    dlamdadt: Sequence[sy.Expr],
    quad_rhs: Sequence[sy.Expr],
) -> GeneratedSource:
    """Emit all problem functions.  ``jac[i, j] = d dydt_i / d states_j``;
    ``dydp[i, k] = d dydt_i / d params[deriv_index[k]]``; ``sens[k, i]`` are symbols."""
    ns, npar, nd = len(states), len(params), len(deriv_index)
    printer = _Printer()
    subs: Dict[sy.Symbol, sy.Symbol] = {time: sy.Symbol('t', real=True)}
    for i, s in enumerate(states):
        subs[s] = sy.Symbol('y[%d]' % i, positive=True)
    for i, s in enumerate(params):
        subs[s] = sy.Symbol('p[%d]' % i, real=True)
    for i, s in enumerate(lamda):
        subs[s] = sy.Symbol('lam[%d]' % i, real=True)
    for k in range(nd):
        for i in range(ns):
            subs[sens[k, i]] = sy.Symbol('s[%d]' % (k * ns + i), real=True)

    T = 'const double t'
    Y = 'const double* __restrict__ y'
    P = 'const double* __restrict__ p'
    LAM = 'const double* __restrict__ lam'
    S = 'const double* __restrict__ s'

    jac = np.asarray(jac, dtype=object).reshape(ns, ns)
    dydp = np.asarray(dydp, dtype=object).reshape(ns, nd)

    col_major = [(i + ns * j) for i in range(ns) for j in range(ns)]
    jac_entries = [jac[i, j] for i in range(ns) for j in range(ns)]
    adj_jac_entries = [-jac[j, i] for i in range(ns) for j in range(ns)]

    # forward sensitivities, row k = d y / d p_k : J s_k + df/dp_k   (out[k*NS + i])
    sens_entries = []
    for k in range(nd):
        for i in range(ns):
            acc = dydp[i, k]
            for j in range(ns):
                acc = acc + jac[i, j] * sens[k, j]
            sens_entries.append(acc)

    all_exprs = list(dydt) + jac_entries + list(quad_rhs)
    parts: List[str] = []
    if _uses_helpers(all_exprs):
        parts.append(_HELPERS)
    parts.append(_emit_function(printer, 'sb_rhs', [T, Y, P], 'out',
                                list(dydt), range(ns), ns, subs, 'f'))
    parts.append(_emit_function(printer, 'sb_jac', [T, Y, P], 'out',
                                jac_entries, col_major, ns * ns, subs, 'j'))
    parts.append(_emit_function(printer, 'sb_adj_rhs', [T, Y, LAM, P], 'out',
                                list(dlamdadt), range(ns), ns, subs, 'a'))
    parts.append(_emit_function(printer, 'sb_adj_jac', [T, Y, P], 'out',
                                adj_jac_entries, col_major, ns * ns, subs, 'b'))
    parts.append(_emit_function(printer, 'sb_quad_rhs', [T, Y, LAM, P], 'out',
                                list(quad_rhs), range(nd), nd, subs, 'q'))
    parts.append(_emit_function(printer, 'sb_sens_rhs', [T, Y, S, P], 'out',
                                sens_entries, range(nd * ns), nd * ns, subs, 's'))
    return GeneratedSource(
        n_states=ns, n_params=npar, n_deriv=nd, body='\n'.join(parts),
        deriv_index=tuple(int(i) for i in deriv_index))
