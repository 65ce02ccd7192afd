"""Generates tests/golden/codegen_golden.npz by running the REFERENCE's own sympy -> numba code
generator (/root/reference/sunode/symode/problem.py + lambdify.py + dtypesubset.py) in this
container.  The reference package cannot be imported as a whole (its __init__ needs the SUNDIALS
cffi extension), so the three pure-Python modules are loaded under a skeleton `sunode` package
with stubs for what they import but never use on this path (`xarray`, `sunode.basic.lib/ffi`,
`sunode.matrix.Sparse`).  Nothing from the reference is copied: the script *runs* it and records
input/output vectors of make_rhs / make_jac_dense / make_adjoint_rhs / make_adjoint_jac_dense /
make_adjoint_quad_rhs on seeded random inputs.

    python tests/golden/make_codegen_golden.py     # needs /root/reference; not run by the tests
"""
import importlib
import importlib.abc
import importlib.util
import os
import sys
import types

import numpy as np

REF = '/root/reference/sunode'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def load_reference():
    pkg = types.ModuleType('sunode')
    pkg.__path__ = [REF]
    sys.modules['sunode'] = pkg
    if 'xarray' not in sys.modules:
        xr = types.ModuleType('xarray')
        xr.DataArray = type('DataArray', (), {})      # only used in isinstance checks
        sys.modules['xarray'] = xr
    basic = types.ModuleType('sunode.basic')
    basic.lib = None
    basic.ffi = None
    basic.data_dtype = np.dtype(np.float64)
    basic.index_dtype = np.dtype(np.int64)
    sys.modules['sunode.basic'] = basic
    pkg.basic = basic
    matrix = types.ModuleType('sunode.matrix')
    matrix.Sparse = object
    sys.modules['sunode.matrix'] = matrix
    sym_pkg = types.ModuleType('sunode.symode')
    sym_pkg.__path__ = [os.path.join(REF, 'symode')]
    sys.modules['sunode.symode'] = sym_pkg

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    pkg.dtypesubset = load('sunode.dtypesubset', os.path.join(REF, 'dtypesubset.py'))
    pkg.problem = load('sunode.problem', os.path.join(REF, 'problem.py'))
    load('sunode.symode.lambdify', os.path.join(REF, 'symode', 'lambdify.py'))
    return load('sunode.symode.problem', os.path.join(REF, 'symode', 'problem.py')).SympyProblem


def main():
    RefProblem = load_reference()
    from tests.golden.problems import CASES     # the same definitions the tests use

    rng = np.random.default_rng(20261017)
    out = {}
    for name, (params, states, rhs, deriv) in CASES.items():
        prob = RefProblem(params, states, rhs, deriv)
        n_s, n_d = prob.n_states, prob.n_params
        # The make_* wrappers themselves do not compile under numba 0.65 (their error branch assigns
        # a record to `user_data.error_states`, which numba can no longer lower).  The generated
        # `compute(_out, ...)` functions they wrap are therefore built exactly as the wrappers
        # build them (reference symode/problem.py:252-258, 285-291, 314-320, 343-349, 407-413).
        lam_mod = sys.modules['sunode.symode.lambdify']

        def gen(tag, argnames, expr):
            calc = lam_mod.lambdify_consts('_%s_%s' % (tag, name), argnames=argnames,
                                           expr=prob._simplify(expr), varmap=prob._varmap)

            def call(out, *args):
                calc(out, *args)
                return int(not np.isfinite(out).all())
            return call

        a3, a4 = ['time', 'state', 'params'], ['time', 'state', 'lamda', 'params']
        rhs_c = gen('rhs', a3, np.array(prob._sym_dydt.T))
        jac_c = gen('jac', a3, prob._sym_dydt_jac)
        adj_c = gen('adj', a4, prob._sym_dlamdadt)
        adjjac_c = gen('adjjac', a3, -prob._sym_dydt_jac.T)
        quad_c = gen('quad', a4, prob._sym_quad_rhs) if n_d else None
        rhs_f = lambda out, t, y, ud: rhs_c(out, t, y, ud.params)
        jac_f = lambda out, t, y, fy, ud: jac_c(out, t, y, ud.params)
        adj_f = lambda out, t, y, lam, ud: adj_c(out, t, y, lam, ud.params)
        adjjac_f = lambda out, t, y, yB, fyB, ud: adjjac_c(out, t, y, ud.params)
        quad_f = lambda out, t, y, lam, ud: quad_c(out, t, y, lam, ud.params)
        n_all = prob.params_dtype.itemsize // 8
        n = 16
        T = rng.uniform(0, 3, n)
        Y = rng.uniform(0.1, 2.0, (n, n_s))
        P = rng.uniform(0.1, 1.5, (n, n_all))
        L = rng.standard_normal((n, n_s))
        R = np.zeros((n, n_s)); J = np.zeros((n, n_s, n_s)); A = np.zeros((n, n_s))
        JB = np.zeros((n, n_s, n_s)); Q = np.zeros((n, n_d))
        for i in range(n):
            ud = prob.make_user_data()
            if n_all:
                ud.params = P[i].view(prob.params_dtype)[0]
            y = Y[i].copy().view(prob.state_dtype)[0]
            assert rhs_f(R[i], T[i], y, ud) == 0
            assert jac_f(J[i], T[i], y, None, ud) == 0
            assert adj_f(A[i], T[i], y, L[i], ud) == 0
            assert adjjac_f(JB[i], T[i], y, None, None, ud) == 0
            if n_d:
                assert quad_f(Q[i], T[i], y, L[i], ud) == 0
        for key, val in dict(t=T, y=Y, p=P, lam=L, rhs=R, jac=J, adj=A, adjjac=JB, quad=Q).items():
            out['%s__%s' % (name, key)] = val
        out['%s__sizes' % name] = np.array([n_s, n_all, n_d])
    # layouts as the reference's DTypeSubset derives them (dtypesubset.py:145-213)
    import json
    layouts = {}
    for name, (params, states, rhs, deriv) in CASES.items():
        prob = RefProblem(params, states, rhs, deriv)
        ps = prob.params_subset
        layouts[name] = {
            'params_dtype': repr(prob.params_dtype), 'state_dtype': repr(prob.state_dtype),
            'subset_dtype': repr(ps.subset_dtype), 'subset_view_dtype': repr(ps.subset_view_dtype),
            'remainder_subset_dtype': repr(ps.remainder.subset_dtype),
            'n_states': prob.n_states, 'n_params': prob.n_params,
            'paths': ['.'.join(p) for p in ps.paths],
            'subset_paths': ['.'.join(p) for p in ps.subset_paths],
            'flat_slices': {'.'.join(k): [v.start, v.stop] for k, v in ps.flat_slices.items()},
            'state_flat_slices': {'.'.join(k): [v.start, v.stop]
                                  for k, v in prob.state_subset.flat_slices.items()},
            'user_data_itemsize': prob.user_data_dtype.itemsize,
        }
    with open(os.path.join(HERE, 'layout_golden.json'), 'w') as fh:
        json.dump(layouts, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, 'codegen_golden.npz'), **out)
    print('wrote', os.path.join(HERE, 'codegen_golden.npz'), len(out), 'arrays')


if __name__ == '__main__':
    main()
