"""Child process of tests/test_pytensor_graph.py -- TEST INFRASTRUCTURE ONLY.

Runs the GRAPH half of sunode_b200/wrappers/as_pytensor.py (solve_ivp, Op.make_node, Op.grad),
which needs PyTensor, on the miniature stand-in of tests/emu/mini_pytensor and on the stand-in
CUDA driver (kernels executed by the host emulation).  The scenario is the reference's own test,
sunode/test_pytensor.py::test_nodiff_params (same problem, same calls), with the closed form
attached instead of shape checks only:  A' = A, B' = B, C' = C  =>  A(t) = A0 e^t."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, 'mini_pytensor'))
sys.path.insert(0, ROOT)

from tests.emu import dryrun_plugin                      # noqa: E402
dryrun_plugin.pytest_configure(None)                     # cubin requests -> emulated kernels

import pytensor                                          # noqa: E402  (the stand-in)
import pytensor.tensor as pt                             # noqa: E402
from sunode_b200.wrappers import as_pytensor             # noqa: E402

assert as_pytensor.HAVE_PYTENSOR and 'mini_pytensor' in pytensor.__file__


def dydt_dict(t, y, p):
    return {'A': y.A, 'B': y.B, 'C': y.C}


A = pt.dscalar("A")
A.tag.test_value = np.array(0.9)
time = pt.linspace(0, 1, 5)
y0 = {'A': (A, ()), 'B': np.array(1.), 'C': np.array(1.)}
beta = pt.dscalar("beta")
params = {'alpha': np.array(1.), 'beta': beta, 'extra': np.array([0.])}
tv = np.linspace(0, 1, 5)

# ---- forward sensitivities (SolveODE; y0 rides along as `__initial_values` pseudo-parameters)
out = as_pytensor.solve_ivp(y0=y0, params=params, rhs=dydt_dict, tvals=time, t0=0.,
                            derivatives="forward", solver_kwargs=dict(sens_mode="simultaneous"))
solution, flat_solution, problem, solver, y0_flat, params_subs_flat, flat_sens, wrapper = out
assert [tuple(p) for p in problem.params_subset.subset_paths] == [('beta',), ('__initial_values', 'A')]
grad_t = pt.grad(solution["A"].sum(), time)
grad_A, grad_beta = pt.grad(solution["A"].sum(), [A, beta])
func = pytensor.function([A, beta], [solution["A"], solution["B"], grad_t, grad_A, grad_beta, flat_sens])
res = func(0.2, 1.)
assert res[0].shape == (5,) and res[2].shape == (5,)                      # the reference's assertions
np.testing.assert_allclose(res[0], 0.2 * np.exp(tv), rtol=1e-8)
np.testing.assert_allclose(res[1], np.exp(tv), rtol=1e-8)
np.testing.assert_allclose(res[2], 0.2 * np.exp(tv), rtol=1e-7)            # d sum_i A(t_i) / d t_i = A'(t_i)
np.testing.assert_allclose(res[3], np.sum(np.exp(tv)), rtol=1e-7)          # d / d A0
np.testing.assert_allclose(res[4], 0.0, atol=1e-9)                         # beta does not act
assert res[5].shape == (5, 2, 3)
print('forward graph ok')

# ---- adjoint (SolveODEAdjoint; its gradient is SolveODEAdjointBackward)
solution, flat_solution, problem, solver, y0_flat, params_subs_flat = as_pytensor.solve_ivp(
    y0=y0, params=params, rhs=dydt_dict, tvals=time, t0=0., derivatives="adjoint")
assert [tuple(p) for p in problem.params_subset.subset_paths] == [('beta',)]
cost = solution["A"].sum() + (solution["B"] * solution["B"]).sum()
grad_t = pt.grad(solution["A"].sum(), time)
grad_A, grad_beta = pt.grad(cost, [A, beta])
func = pytensor.function([A, beta], [solution["A"], solution["B"], grad_t, grad_A, grad_beta])
res = func(0.2, 1.)
assert res[0].shape == (5,) and res[2].shape == (5,)                      # the reference's assertions
np.testing.assert_allclose(res[0], 0.2 * np.exp(tv), rtol=1e-8)
np.testing.assert_allclose(res[2], 0.2 * np.exp(tv), rtol=1e-7)
np.testing.assert_allclose(res[3], np.sum(np.exp(tv)), rtol=1e-7)          # through -lamda(t0) and y0_flat
np.testing.assert_allclose(res[4], 0.0, atol=1e-8)
# a failing solve gives NaN outputs, not an exception (as_pytensor.py:287-290)
res = pytensor.function([A, beta], [solution["A"], grad_A])(np.inf, 1.)
assert np.isnan(res[0]).all() and np.isnan(res[1]).all()
print('adjoint graph ok')

# ---- the batched Ops in a graph: B draws per call, gradients through SolveODEAdjointBackwardBatch
from sunode_b200 import SympyProblem                       # noqa: E402
from sunode_b200.solver import AdjointSolver               # noqa: E402
bprob = SympyProblem({'a': {'b': ()}}, {'x': ()}, lambda t, y, p: {'x': y.x + p.a.b}, [('a', 'b')])
bsolver = AdjointSolver(bprob)
Y0 = pt.dmatrix("Y0")
P = pt.dmatrix("P")
yb = as_pytensor.SolveODEAdjointBatch(bsolver)(Y0, P, np.zeros(0), 0.0, tv)
gy0, gp = pt.grad(yb.sum(), [Y0, P])
fb = pytensor.function([Y0, P], [yb, gy0, gp])
x0, b = np.array([[1.0], [0.5], [2.0]]), np.array([[0.2], [0.1], [-0.3]])
yv, g_y0, g_p = fb(x0, b)
np.testing.assert_allclose(yv[:, :, 0], (x0 + b) * np.exp(tv) - b, rtol=1e-8)
np.testing.assert_allclose(g_y0[:, 0], np.full(3, np.sum(np.exp(tv))), rtol=1e-7)
np.testing.assert_allclose(g_p[:, 0], np.full(3, np.sum(np.exp(tv) - 1)), rtol=1e-7)
print('batched graph ok')

# ---- a wrong-rank input is refused when the node is built (itypes)
try:
    as_pytensor.SolveODEAdjoint(solver)(np.zeros((2, 2)), np.zeros(1), np.zeros(2), 0.0, tv)
    raise AssertionError('expected TypeError')
except TypeError:
    pass
print('ALL OK')
