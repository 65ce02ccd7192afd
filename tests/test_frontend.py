"""Front-end behaviour the reference's smoke tests pin (sunode/test_solve.py:7-78): problems with
one parameter, no parameters, nested parameters and nested states construct; plus the error
behaviour of SympyProblem (symode/problem.py:211-228) and the parameter plumbing
(symode/problem.py:232-243)."""
import pickle

import numpy as np
import pytest
import sympy as sym

from sunode_b200 import SympyProblem


def test_nodiff_params():
    def rhs(t, y, p):
        return {'x': y.x}

    prob = SympyProblem({'b': ()}, {'x': ()}, rhs, [])
    assert prob.n_states == 1 and prob.n_params == 0 and prob.n_params_total == 1
    assert prob.generated.n_deriv == 0


def test_empty_params():
    def rhs(t, y, p):
        return {'x': y.x}

    prob = SympyProblem({}, {'x': ()}, rhs, [])
    assert prob.params_dtype.itemsize == 0 and prob.n_params_total == 0
    out = np.zeros(1)
    assert prob.make_rhs()(out, 0.0, np.array([2.0]), prob.make_user_data()) == 0
    assert out[0] == 2.0


def test_nested_params_and_states():
    def rhs(t, y, p):
        return {'a': {'x': y.a.x * p.a.b, 'y': [y.a.y[0], p.c * y.a.y[1]]}}

    prob = SympyProblem({'a': {'b': ()}, 'c': ()}, {'a': {'x': (), 'y': 2}}, rhs, [('a', 'b')])
    assert prob.n_states == 3 and prob.n_params == 1 and prob.n_params_total == 2
    assert prob.state_dtype['a']['y'].shape == (2,)
    ud = prob.make_user_data()
    prob.update_subset_params(ud, np.array([(3.0,)], dtype=prob.params_subset.subset_dtype)[0])
    prob.update_remaining_params(ud, np.array([(5.0,)], dtype=prob.params_subset.remainder.subset_dtype)[0])
    assert ud.params.a.b == 3.0 and ud.params.c == 5.0
    np.testing.assert_array_equal(prob.flat_params(ud), [3.0, 5.0])
    out = np.zeros(3)
    assert prob.make_rhs()(out, 0.0, np.array([1.0, 2.0, 4.0]), ud) == 0
    np.testing.assert_array_equal(out, [3.0, 2.0, 20.0])
    sol = prob.flat_solution_as_dict(np.arange(6.0).reshape(2, 3))
    np.testing.assert_array_equal(sol['a']['y'], [[1, 2], [4, 5]])


def test_coords_and_dict_valued_rhs():
    def rhs(t, y, p):
        return {'c': {'u': p.k * y.c[0], 'v': -y.c[1]}}

    prob = SympyProblem({'k': ()}, {'c': 'species'}, rhs, [('k',)], coords={'species': ['u', 'v']})
    assert prob.n_states == 2


def test_missing_and_unknown_states_raise():
    with pytest.raises(ValueError, match='No right-hand-side'):
        SympyProblem({}, {'x': (), 'y': ()}, lambda t, y, p: {'x': y.x}, [])
    with pytest.raises(ValueError, match='Unknown state'):
        SympyProblem({}, {'x': ()}, lambda t, y, p: {'x': y.x, 'z': y.x}, [])
    with pytest.raises(ValueError, match='shape'):
        SympyProblem({}, {'x': 2}, lambda t, y, p: {'x': [y.x[0]]}, [])
    with pytest.raises(ValueError):
        SympyProblem({'a': ()}, {'x': ()}, lambda t, y, p: {'x': y.x}, [('nope',)])


def test_symbol_assumptions_and_derived_expressions():
    """states positive, params real (symode/problem.py:78-79); J, f_p, -lam^T J, lam^T f_p."""
    def rhs(t, y, p):
        return {'x': p.a * sym.sqrt(y.x ** 2) + sym.sin(t)}

    prob = SympyProblem({'a': ()}, {'x': ()}, rhs, [('a',)])
    x, = prob._sym_statevec
    a, = prob._sym_paramsvec
    lam, = prob._sym_lamda
    assert x.is_positive and a.is_real
    assert prob._sym_dydt_jac[0, 0] == a                  # sqrt(x**2) -> x because x > 0
    assert prob._sym_dlamdadt[0] == -lam * a
    assert prob._sym_quad_rhs[0] == lam * x


def test_problem_pickles():
    def rhs(t, y, p):
        return {'x': -p.k * y.x}

    import types
    mod = types.ModuleType('_pickle_rhs_mod')
    mod.rhs = rhs
    rhs.__module__ = '_pickle_rhs_mod'
    rhs.__qualname__ = 'rhs'
    import sys
    sys.modules['_pickle_rhs_mod'] = mod
    prob = SympyProblem({'k': ()}, {'x': ()}, rhs, [('k',)])
    digest = prob.generated.digest
    clone = pickle.loads(pickle.dumps(prob))
    assert clone.generated.digest == digest
