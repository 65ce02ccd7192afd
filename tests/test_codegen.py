"""Code generator + layouts against golden vectors produced by RUNNING the reference's own
sympy -> numba generator and DTypeSubset (tests/golden/make_codegen_golden.py).  Tolerance: the
reference compiles with fastmath (lambdify.py:88) and a different CSE, so agreement is to
rounding level (1e-13 relative), not bit-exact."""
import json
import os

import numpy as np
import pytest

from sunode_b200 import SympyProblem
from tests.golden.problems import CASES

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'codegen_golden.npz'))
with open(os.path.join(HERE, 'golden', 'layout_golden.json')) as fh:
    LAYOUT = json.load(fh)

# (rtol, atol).  'helpers' sums O(1) terms that cancel and goes through exp / log1p, where numba's
# fastmath intrinsics and libm differ by an ulp: rounding level is 1e-13 absolute there.
TOL = {name: (1e-13, 1e-14) for name in CASES}
TOL['helpers'] = (1e-12, 2e-13)


@pytest.mark.parametrize('name', list(CASES))
def test_generated_functions_match_reference_generator(name):
    params, states, rhs, deriv = CASES[name]
    prob = SympyProblem(params, states, rhs, deriv)
    rtol, atol = TOL[name]
    host = prob.host_functions
    n_s, n_all, n_d = GOLD['%s__sizes' % name]
    assert (prob.n_states, prob.n_params_total, prob.n_params) == (n_s, n_all, n_d)
    T, Y, P, L = (GOLD['%s__%s' % (name, k)] for k in ('t', 'y', 'p', 'lam'))
    for i in range(len(T)):
        out = np.zeros(n_s)
        assert host.rhs(T[i], Y[i], P[i], out) == 0
        np.testing.assert_allclose(out, GOLD[name + '__rhs'][i], rtol=rtol, atol=max(atol, 1e-15))
        J = np.zeros(n_s * n_s)
        assert host.jac(T[i], Y[i], P[i], J) == 0
        # ours is column-major (the layout the reference hands to SUNDenseMatrix, problem.py:345)
        np.testing.assert_allclose(J.reshape(n_s, n_s).T, GOLD[name + '__jac'][i], rtol=rtol, atol=max(atol, 1e-15))
        assert host.adj_rhs(T[i], Y[i], L[i], P[i], out) == 0
        np.testing.assert_allclose(out, GOLD[name + '__adj'][i], rtol=rtol, atol=max(atol, 1e-15))
        assert host.adj_jac(T[i], Y[i], P[i], J) == 0
        np.testing.assert_allclose(J.reshape(n_s, n_s).T, GOLD[name + '__adjjac'][i], rtol=rtol, atol=max(atol, 1e-15))
        if n_d:
            q = np.zeros(n_d)
            assert host.quad_rhs(T[i], Y[i], L[i], P[i], q) == 0
            np.testing.assert_allclose(q, GOLD[name + '__quad'][i], rtol=rtol, atol=max(atol, 1e-15))


@pytest.mark.parametrize('name', list(CASES))
def test_python_callables_keep_reference_signatures(name):
    """make_rhs()/make_jac_dense()/... are called as in the reference
    (symode/problem.py:262,353,294,323,417; as_pytensor.py:173-178)."""
    params, states, rhs, deriv = CASES[name]
    prob = SympyProblem(params, states, rhs, deriv)
    rtol, atol = TOL[name]
    n_s, n_all, n_d = GOLD['%s__sizes' % name]
    T, Y, P, L = (GOLD['%s__%s' % (name, k)] for k in ('t', 'y', 'p', 'lam'))
    ud = prob.make_user_data()
    ud.params = P[0].copy().view(prob.params_dtype)[0]
    y = Y[0].copy().view(prob.state_dtype)[0]
    out = np.zeros(n_s)
    assert prob.make_rhs()(out, T[0], y, ud) == 0
    np.testing.assert_allclose(out, GOLD[name + '__rhs'][0], rtol=rtol, atol=max(atol, 1e-15))
    J = np.zeros((n_s, n_s))
    assert prob.make_jac_dense()(J, T[0], y, None, ud) == 0
    np.testing.assert_allclose(J, GOLD[name + '__jac'][0], rtol=rtol, atol=max(atol, 1e-15))
    assert prob.make_adjoint_rhs()(out, T[0], y, L[0], ud) == 0
    np.testing.assert_allclose(out, GOLD[name + '__adj'][0], rtol=rtol, atol=max(atol, 1e-15))
    assert prob.make_adjoint_jac_dense()(J, T[0], y, None, None, ud) == 0
    np.testing.assert_allclose(J, GOLD[name + '__adjjac'][0], rtol=rtol, atol=max(atol, 1e-15))
    q = np.zeros(n_d)
    assert prob.make_adjoint_quad_rhs()(q, T[0], y, L[0], ud) == 0
    np.testing.assert_allclose(q, GOLD[name + '__quad'][0], rtol=rtol, atol=max(atol, 1e-15))
    # J v and -J^T v products, sensitivity rhs  (symode/problem.py:373-465, 557-583)
    v = L[0]
    jv = np.zeros(n_s)
    assert prob.make_rhs_jac_prod()(jv, v, T[0], y, None, ud) == 0
    np.testing.assert_allclose(jv, GOLD[name + '__jac'][0] @ v, rtol=1e-12, atol=1e-14)
    assert prob.make_adjoint_jac_prod()(jv, v, T[0], y, None, None, ud) == 0
    np.testing.assert_allclose(jv, GOLD[name + '__adjjac'][0] @ v, rtol=1e-12, atol=1e-14)


def test_nonfinite_output_returns_one():
    """Reference callbacks return 1 ("recoverable") on non-finite output and stash the state
    (symode/problem.py:266-270)."""
    params, states, rhs, deriv = CASES['one_fixed']
    prob = SympyProblem(params, states, rhs, deriv)
    ud = prob.make_user_data()
    ud.params = np.array([1.0, 0.0]).view(prob.params_dtype)[0]      # K = 0 -> division by zero
    out = np.zeros(1)
    assert prob.make_rhs()(out, 0.0, np.array([0.5]), ud) == 1
    assert not np.isfinite(ud.error_rhs).all()


@pytest.mark.parametrize('name', list(CASES))
def test_layouts_match_reference(name):
    params, states, rhs, deriv = CASES[name]
    prob = SympyProblem(params, states, rhs, deriv)
    g = LAYOUT[name]
    ps = prob.params_subset
    assert repr(prob.params_dtype) == g['params_dtype']
    assert repr(prob.state_dtype) == g['state_dtype']
    assert repr(ps.subset_dtype) == g['subset_dtype']
    assert repr(ps.subset_view_dtype) == g['subset_view_dtype']
    assert repr(ps.remainder.subset_dtype) == g['remainder_subset_dtype']
    assert prob.n_states == g['n_states'] and prob.n_params == g['n_params']
    assert ['.'.join(p) for p in ps.paths] == g['paths']
    assert ['.'.join(p) for p in ps.subset_paths] == g['subset_paths']
    assert {'.'.join(k): [v.start, v.stop] for k, v in ps.flat_slices.items()} == g['flat_slices']
    assert {'.'.join(k): [v.start, v.stop]
            for k, v in prob.state_subset.flat_slices.items()} == g['state_flat_slices']
    assert prob.user_data_dtype.itemsize == g['user_data_itemsize']


def test_cuda_and_c_flavours_share_bodies():
    prob = SympyProblem(*CASES['lv'][:3], CASES['lv'][3])
    gen = prob.generated
    assert '__device__' in gen.cuda and '__device__' not in gen.c
    for fn in ('sb_rhs', 'sb_jac', 'sb_adj_rhs', 'sb_adj_jac', 'sb_quad_rhs', 'sb_sens_rhs'):
        assert fn in gen.cuda and fn in gen.c
    assert gen.deriv_index == (0, 1)
