"""Generates tests/golden/rewrite_golden.json by running the REFERENCE's log-sum-exp rewrites
(/root/reference/sunode/symode/lambdify.py:355-432: logsumexp_2terms_opt, explog_opt) on a list of
expressions and evaluating the rewritten expressions at seeded points.

    python tests/golden/make_rewrite_golden.py     # needs /root/reference; not run by the tests
"""
import json
import os
import sys

import numpy as np
import sympy as sy
import sympy.codegen.rewriting as rw

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

SYMS = sy.symbols('c1 c2 c3', real=True)


def expressions():
    c1, c2, c3 = SYMS
    return {
        'softmax2': sy.exp(c2) / (sy.exp(c1) + sy.exp(c2)),
        'softmax2_sq': sy.exp(c2) / 2 / (sy.exp(c1) + sy.exp(c2)) ** 2,
        'logsum': sy.log(sy.exp(c1) + sy.exp(c2)),
        'inside_sum': c3 * sy.exp(c1) / (sy.exp(c1) + sy.exp(c2)) + 1,
        'negative': -sy.exp(c1) / (sy.exp(c1) + sy.exp(c3)),
        'untouched': c1 * sy.exp(c2) + c3,
    }


def evaluate(expr, points, logaddexp_cls):
    f = sy.lambdify(SYMS, expr, modules=[{logaddexp_cls.__name__: np.logaddexp}, 'numpy'])
    return [float(f(*p)) for p in points]


def main():
    from tests.golden.make_codegen_golden import load_reference
    load_reference()
    lam = sys.modules['sunode.symode.lambdify']
    rng = np.random.default_rng(20261017)
    points = rng.uniform(-3, 3, (8, 3)).tolist() + [[700.0, 705.0, -2.0], [-720.0, -715.0, 1.5]]
    out = {'points': points, 'cases': {}}
    for name, e in expressions().items():
        case = {}
        for tag, opt in (('explog', lam.explog_opt), ('logsumexp', lam.logsumexp_2terms_opt)):
            r = rw.optimize(e, [opt])
            with np.errstate(all='ignore'):
                case[tag] = {'uses_logaddexp': bool(r.atoms(lam.logaddexp)), 'changed': r != e,
                             'values': evaluate(r, points, lam.logaddexp)}
        out['cases'][name] = case
    with open(os.path.join(HERE, 'rewrite_golden.json'), 'w') as fh:
        json.dump(out, fh, indent=1)
    print('wrote rewrite_golden.json', {k: (v['explog']['changed'], v['logsumexp']['changed']) for k, v in out['cases'].items()})


if __name__ == '__main__':
    main()
