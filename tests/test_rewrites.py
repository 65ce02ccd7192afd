"""The log-sum-exp rewrites (sunode_b200/symode/functions.py) against goldens recorded from the
reference's own ``explog_opt`` / ``logsumexp_2terms_opt`` (tests/golden/make_rewrite_golden.py):
the same expressions are changed, `logaddexp` appears in the same ones, and the rewritten
expressions take the same values -- including at points where the naive form overflows."""
import json
import os

import numpy as np
import pytest
import sympy as sy
import sympy.codegen.rewriting as rw

from sunode_b200 import SympyProblem
from sunode_b200.symode import functions as fn
from tests.golden.make_rewrite_golden import SYMS, evaluate, expressions

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, 'golden', 'rewrite_golden.json')) as fh:
    GOLD = json.load(fh)


@pytest.mark.parametrize('name', list(expressions()))
@pytest.mark.parametrize('tag', ['explog', 'logsumexp'])
def test_rewrite_matches_reference(name, tag):
    e = expressions()[name]
    opt = fn.explog_opt if tag == 'explog' else fn.logsumexp_2terms_opt
    r = rw.optimize(e, [opt])
    g = GOLD['cases'][name][tag]
    assert (r != e) == g['changed']
    assert bool(r.atoms(fn.logaddexp)) == g['uses_logaddexp']
    with np.errstate(all='ignore'):
        vals = np.array(evaluate(r, GOLD['points'], fn.logaddexp))
    ref = np.array(g['values'], dtype=float)
    np.testing.assert_array_equal(np.isfinite(vals), np.isfinite(ref))
    ok = np.isfinite(ref)
    np.testing.assert_allclose(vals[ok], ref[ok], rtol=1e-13, atol=1e-300)


def test_rewritten_problem_compiles_and_agrees():
    """``simplify=`` with the rewrite, as a user of the reference would pass it: the generated
    functions contain ``sb_logaddexp`` and agree with the plain problem to rounding."""
    def rhs(t, y, p):
        return {'u': sy.exp(p.a) / (sy.exp(p.a) + sy.exp(y.u)) - y.u * sy.log(sy.exp(p.b) + sy.exp(y.u))}

    spec = ({'a': (), 'b': ()}, {'u': ()}, rhs, [('a',), ('b',)])
    plain = SympyProblem(*spec)
    stable = SympyProblem(*spec, simplify=lambda e: rw.optimize(e, [fn.explog_opt, fn.logsumexp_2terms_opt]))
    assert 'sb_logaddexp' in stable.generated.cuda and 'sb_logaddexp' not in plain.generated.cuda
    rng = np.random.default_rng(3)
    for _ in range(8):
        y, p, lam = rng.uniform(0.1, 2, 1), rng.uniform(-2, 2, 2), rng.standard_normal(1)
        for name, n_out, args in (('rhs', 1, (y, p)), ('jac', 1, (y, p)), ('adj_rhs', 1, (y, lam, p)),
                                  ('quad_rhs', 2, (y, lam, p))):
            a, b = np.zeros(n_out), np.zeros(n_out)
            assert getattr(plain.host_functions, name)(0.3, *args, a) == 0
            assert getattr(stable.host_functions, name)(0.3, *args, b) == 0
            np.testing.assert_allclose(b, a, rtol=1e-12, atol=1e-14)
    # where the naive form overflows the rewritten one does not
    a, b = np.zeros(1), np.zeros(1)
    y, p = np.array([0.5]), np.array([800.0, 1.0])
    plain.host_functions.rhs(0.0, y, p, a)
    assert stable.host_functions.rhs(0.0, y, p, b) == 0
    assert not np.isfinite(a[0]) and np.isfinite(b[0])
