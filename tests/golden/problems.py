"""Problem definitions shared by the golden-vector generator (which feeds them to the REFERENCE's
SympyProblem) and the tests (which feed them to ours).  name -> (params, states, rhs, deriv)."""
import numpy as np
import sympy as sym


def _lv(t, y, p):
    return {'hares': p.alpha * y.hares - p.beta * y.lynx * y.hares,
            'lynx': p.delta * y.hares * y.lynx - p.gamma * y.lynx}


def _robertson(t, y, p):
    return {'y1': -p.k1 * y.y1 + p.k3 * y.y2 * y.y3,
            'y2': p.k1 * y.y1 - p.k3 * y.y2 * y.y3 - p.k2 * y.y2 ** 2,
            'y3': p.k2 * y.y2 ** 2}


def _nested(t, y, p):
    # nested params/states, a vector state, time dependence and transcendental functions
    return {
        'a': p.c.d * y.a + p.f[2] * sym.sin(t),
        'b': {'c': [3. * sym.exp(-y.b.c[1]), 4. * y.a[0] / (1 + y.b.c[0] ** 2)]},
    }


def _one_fixed(t, y, p):
    # derivative wrt the second parameter only; the first one is a "remaining" parameter
    return {'x': p.r * y.x * (1 - y.x / p.K) + sym.sqrt(y.x)}


def _seir(t, y, p):
    # BASELINE configs[4] (sunode_b200/examples.py: two-group SEIR, nested states, all six
    # parameters differentiated)
    out = {}
    betas = (p.beta1, p.beta2)
    groups = (y.g1, y.g2)
    for k, name in enumerate(('g1', 'g2')):
        g, h = groups[k], groups[1 - k]
        lam = betas[k] * g.I + p.kappa * h.I
        out[name] = {'S': -g.S * lam + p.omega * g.R, 'E': g.S * lam - p.sigma * g.E,
                     'I': p.sigma * g.E - p.gamma * g.I, 'R': p.gamma * g.I - p.omega * g.R}
    return out


def _helpers(t, y, p):
    # the helper functions the reference makes available inside generated code
    # (sunode/symode/lambdify.py:59-77, 275-352): logaddexp is differentiated wrt a state and a
    # parameter (its fdiff works upstream); expit / dexpit / the cardinal B-spline only take the
    # time as argument, because the reference's expit.fdiff / dexpit.fdiff raise when sympy
    # differentiates through them (lambdify.py:301-305, 318-322) and CardinalBSpline has no fdiff
    import sys
    fn = sys.modules.get('sunode.symode.lambdify')          # the generator run: the reference's
    if fn is None:
        from sunode_b200.symode import functions as fn      # the tests: ours
    forcing = fn.interpolate_spline(t, [p.c[0], p.c[1], p.c[2], 0.5, 0.25], 0, 3, 4)
    return {'u': -p.k * fn.expit(t - 1) * y.u + forcing + fn.logaddexp(y.v, p.k),
            'v': fn.dexpit(2 * t - 1) * y.u - y.v * fn.logaddexp(p.c[0], y.u)}


_SEIR_PARAMS = ['beta1', 'beta2', 'kappa', 'sigma', 'gamma', 'omega']
_SEIR_GROUP = {'S': (), 'E': (), 'I': (), 'R': ()}

CASES = {
    'lv': ({'alpha': (), 'beta': (), 'gamma': (), 'delta': ()}, {'hares': (), 'lynx': ()}, _lv,
           [('alpha',), ('beta',)]),
    'robertson': ({'k1': (), 'k2': (), 'k3': ()}, {'y1': (), 'y2': (), 'y3': ()}, _robertson,
                  [('k1',), ('k2',), ('k3',)]),
    'nested': ({'c': {'d': 3}, 'f': 4}, {'a': 3, 'b': {'c': 2}}, _nested, [('c', 'd')]),
    'one_fixed': ({'r': (), 'K': ()}, {'x': ()}, _one_fixed, [('K',)]),
    # appended last: the generator draws its inputs sequentially, earlier cases keep their vectors
    'seir': ({n: () for n in _SEIR_PARAMS}, {'g1': dict(_SEIR_GROUP), 'g2': dict(_SEIR_GROUP)}, _seir,
             [(n,) for n in _SEIR_PARAMS]),
    'helpers': ({'k': (), 'c': 3}, {'u': (), 'v': ()}, _helpers, [('k',), ('c',)]),
}
