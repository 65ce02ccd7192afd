"""Reference-held goldens through the CUDA path (``-m gpu``).

tests/test_oracle.py pins the CPU oracle to what the reference ships; the tests below put the
*kernels* against the same vectors directly, so that the chain CUDA == CVODES does not lean on the
oracle in between:

  G1  the CVODES run recorded in notebooks/from_sympy.ipynb:240-242 (5 states -> the lane-group
      backward kernel, a fixed parameter block of 50, gradients wrt y0 and the parameters);
  G2  the double integrator of from_sympy.ipynb cells 39-41 against its closed form (every error
      estimate is exactly zero from order 2 on: the controller's zero-error branch on the device);
  G3  the closed form of the reference's smoke-test problem (sunode/test_solve.py:81-154);
  codegen  ``sb_eval`` (the device flavour of the generated functions) against the vectors the
      reference's own sympy -> numba generator produced (tests/golden/make_codegen_golden.py),
      including the helper functions logaddexp / expit / dexpit / CardinalBSpline;
  Robertson / SEIR  against SciPy Radau at 1e-12 (nothing shared with BDF), relative to the
      oracle's own distance from that truth.
"""
import os

import numpy as np
import pytest

from sunode_b200 import SympyProblem, examples
from sunode_b200.solver import AdjointSolver, Solver
from tests.golden.problems import CASES

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'codegen_golden.npz'))


def _g1_problem():
    def rhs(t, y, p):
        return {'a': p.c.d * y.a + p.f[20], 'b': {'c': [3., 4.]}}

    return SympyProblem(params={'c': {'d': 3}, 'f': 50}, states={'a': 3, 'b': {'c': 2}},
                        rhs_sympy=rhs, derivative_params=[('c', 'd')])


@pytest.mark.parametrize('tol,rtol_val,rtol_grad', [(1e-10, 5e-9, 2e-8), (1e-12, 3e-11, 1e-9)])
def test_g1_notebook_cvodes_run_on_the_gpu(tol, rtol_val, rtol_grad):
    """Same inputs, same assertions and tolerances as tests/test_oracle.py::test_g1 -- but the
    numbers come from sb_forward / sb_tables / sb_backward, through the reference-shaped
    batch-1 API (solve_forward + solve_backward, from_sympy.ipynb:178-179)."""
    rs = np.random.RandomState(42)          # np.random.seed(42); b = randn(2); d = randn(3)
    b, d = rs.randn(2), rs.randn(3)
    f = np.linspace(0, 1, 50)
    prob = _g1_problem()
    assert prob.n_states == 5 and prob.n_params == 3
    tvals = np.arange(20) / 100
    y0 = np.concatenate([np.arange(3, dtype=float) + d[0] ** 2, b ** 3])
    solver = AdjointSolver(prob, abstol=tol, reltol=tol)
    solver.set_params_dict({'c': {'d': d}, 'f': f})
    y_out, grad_out, lamda_out = solver.make_output_buffers(tvals)
    solver.solve_forward(0, tvals, y0, y_out)
    val = np.sum(y_out ** 2)
    solver.solve_backward(tvals[-1], 0, tvals, 2 * y_out, grad_out, lamda_out)
    dy0 = -lamda_out                         # as_pytensor.py:303
    grad_d = grad_out.copy()
    grad_d[0] += np.sum(dy0[:3]) * 2 * d[0]  # y0.a = arange(3) + d[0]**2
    grad_b = dy0[3:] * 3 * b ** 2            # y0.b.c = b**3
    np.testing.assert_allclose(val, 185.95454144, rtol=rtol_val)
    np.testing.assert_allclose(grad_b, [12.06638293, 0.86567236], rtol=rtol_grad)
    np.testing.assert_allclose(grad_d, [252.23687613, 12.10402814, 21.63579496], rtol=rtol_grad)


def test_g1_batched_and_batch1_agree():
    """The same G1 instance inside a batch of perturbed ones (device layout, grouped lanes with
    padding instances in the last warp) gives bit-identical numbers to the batch-1 call."""
    rs = np.random.RandomState(42)
    b, d = rs.randn(2), rs.randn(3)
    f = np.linspace(0, 1, 50)
    prob = _g1_problem()
    tvals = np.arange(20) / 100
    y0 = np.concatenate([np.arange(3, dtype=float) + d[0] ** 2, b ** 3])
    params = np.concatenate([d, f])
    rng = np.random.default_rng(5)
    B = 37
    Y0 = y0 * (1 + 0.05 * rng.standard_normal((B, 5)))
    P = params * (1 + 0.05 * rng.standard_normal((B, 53)))
    Y0[11], P[11] = y0, params
    solver = AdjointSolver(prob, abstol=1e-10, reltol=1e-10)
    y1, st1 = solver.solve_forward_batch(0.0, tvals, y0[None], params[None])
    g1, l1, sb1 = solver.solve_backward_batch(tvals[-1], 0.0, tvals, 2 * y1)
    yB, stB = solver.solve_forward_batch(0.0, tvals, Y0, P)
    gB, lB, sbB = solver.solve_backward_batch(tvals[-1], 0.0, tvals, 2 * yB)
    assert (stB == 0).all() and (sbB == 0).all() and st1[0] == 0 and sb1[0] == 0
    np.testing.assert_array_equal(yB[11], y1[0])
    np.testing.assert_array_equal(gB[11], g1[0])
    np.testing.assert_array_equal(lB[11], l1[0])


def _double_integrator():
    return SympyProblem(params={'a': (), 'b': (), 'c': ()}, states={'x': (), 'v': ()},
                        rhs_sympy=lambda t, y, p: {'x': y.v, 'v': p.b},
                        derivative_params=[('a',), ('b',), ('c',)])


def test_g2_double_integrator_closed_form_on_the_gpu():
    """from_sympy.ipynb cells 39-41 record that the adjoint loss / gradients equal the analytic
    ones to ~1e-11 relative; same envelope as tests/test_oracle.py::test_g2 (1e-9 / 1e-8)."""
    prob = _double_integrator()
    tvals = np.arange(1, 10).astype(float)
    rng = np.random.default_rng(41)
    B = 64
    p, y0 = rng.standard_normal((B, 3)), rng.standard_normal((B, 2))
    x = 0.5 * tvals ** 2 * p[:, 1:2] + tvals * y0[:, 1:2] + y0[:, 0:1]
    v = tvals * p[:, 1:2] + y0[:, 1:2]
    sol = np.stack([x, v], axis=2)
    solver = AdjointSolver(prob, abstol=1e-10, reltol=1e-10)
    y, st = solver.solve_forward_batch(0.0, tvals, y0, p)
    assert (st == 0).all()
    np.testing.assert_allclose(y, sol, rtol=1e-9, atol=1e-9)
    grad, lam, sb = solver.solve_backward_batch(tvals[-1], 0.0, tvals, 2 * y)
    assert (sb == 0).all()
    grad_b = np.sum(2 * x * 0.5 * tvals ** 2 + 2 * v * tvals, axis=1)
    grad_y0 = np.stack([np.sum(2 * x, axis=1), np.sum(2 * x * tvals + 2 * v, axis=1)], axis=1)
    scale = np.abs(grad_b).max()
    np.testing.assert_allclose(grad[:, 1], grad_b, rtol=1e-8, atol=1e-8 * scale)
    np.testing.assert_allclose(grad[:, [0, 2]], 0.0, atol=1e-8 * scale)
    np.testing.assert_allclose(-lam, grad_y0, rtol=1e-8, atol=1e-8 * np.abs(grad_y0).max())


def test_g3_smoke_problem_closed_form_on_the_gpu():
    """sunode/test_solve.py:81-154 (x' = x + b) at the reference's default tolerances, with the
    assertions of tests/test_oracle.py::test_g3."""
    prob = SympyProblem({'a': {'b': ()}}, {'x': ()}, lambda t, y, p: {'x': y.x + p.a.b}, [('a', 'b')])
    b, time = 0.2, np.linspace(0, 1)
    solver = AdjointSolver(prob)
    y, grad, lam, status = solver.solve_adjoint_batch(0.0, time, np.ones((1, 1)), np.array([[b]]),
                                                      np.ones((50, 1)))
    assert status[0] == 0
    np.testing.assert_allclose(y[0, :, 0], (1 + b) * np.exp(time) - b, rtol=1e-8)
    np.testing.assert_allclose(grad[0, 0], np.sum(np.exp(time) - 1), rtol=1e-7)
    np.testing.assert_allclose(-lam[0, 0], np.sum(np.exp(time)), rtol=1e-7)


@pytest.mark.parametrize('name', list(CASES))
def test_sb_eval_matches_reference_generator(name):
    """The CUDA flavour of every generated function, evaluated on the device by ``sb_eval``,
    against the vectors recorded from the reference's generator (1e-13 like the host flavour in
    tests/test_codegen.py; 'helpers' goes through exp / log1p and cancelling O(1) sums)."""
    params, states, rhs, deriv = CASES[name]
    prob = SympyProblem(params, states, rhs, deriv)
    rtol, atol = (1e-12, 2e-13) if name == 'helpers' else (1e-13, 1e-14)
    n_s, n_all, n_d = (int(v) for v in GOLD['%s__sizes' % name])
    T, Y, P, L = (np.ascontiguousarray(GOLD['%s__%s' % (name, k)]) for k in ('t', 'y', 'p', 'lam'))
    n = len(T)
    eng = Solver(prob)._engine

    def ev(kind, n_out, lam=None):
        out = np.full((n, n_out), np.nan)
        eng.eval(kind, T, Y, P, lam, out)
        return out

    np.testing.assert_allclose(ev(0, n_s), GOLD[name + '__rhs'], rtol=rtol, atol=atol)
    # column-major on the device (the layout the reference hands to SUNDenseMatrix, problem.py:345)
    jac = ev(1, n_s * n_s).reshape(n, n_s, n_s).transpose(0, 2, 1)
    np.testing.assert_allclose(jac, GOLD[name + '__jac'], rtol=rtol, atol=atol)
    np.testing.assert_allclose(ev(2, n_s, L), GOLD[name + '__adj'], rtol=rtol, atol=atol)
    adjjac = ev(4, n_s * n_s).reshape(n, n_s, n_s).transpose(0, 2, 1)
    np.testing.assert_allclose(adjjac, GOLD[name + '__adjjac'], rtol=rtol, atol=atol)
    if n_d:
        np.testing.assert_allclose(ev(3, n_d, L), GOLD[name + '__quad'], rtol=rtol, atol=atol)


def _radau_truth(w, prob, y0, theta):
    from scipy.integrate import solve_ivp
    gen, n_s = prob.host_functions, prob.n_states

    def f(t, y):
        out = np.zeros(n_s)
        gen.rhs(t, np.ascontiguousarray(y), theta, out)
        return out

    def jac(t, y):
        J = np.zeros(n_s * n_s)
        gen.jac(t, np.ascontiguousarray(y), theta, J)
        return J.reshape(n_s, n_s).T
    sol = solve_ivp(f, (w.t0, w.tvals[-1]), y0, method='Radau', jac=jac, t_eval=w.tvals,
                    rtol=1e-12, atol=1e-14)
    assert sol.success
    return sol.y.T


@pytest.mark.parametrize('name,env', [('robertson_adj', 500.0), ('seir_adj', 200.0), ('lv_adj', 300.0)])
def test_gpu_error_against_independent_truth(name, env):
    """GPU and oracle against SciPy Radau at 1e-12 on the same draws.  For the stiff Robertson
    problem the two BDF step sequences decorrelate at rounding level, so |y_gpu - y_oracle| is of
    the order of the global error; what must hold is that the GPU is as close to the truth as
    the oracle is: err_gpu <= 2 err_oracle per draw (floor: 5 tolerance units), and inside the
    envelope tests/test_oracle.py::test_g5 states for the oracle."""
    from oracle.oracle import Oracle
    w = examples.workloads()[name]
    prob = w.make_problem()
    B = 12
    y0, theta = w.draws(B)
    y0[0], theta[0] = np.asarray(w.y0, float), np.asarray(w.theta_med, float)
    y, status = Solver(prob, abstol=1e-8, reltol=1e-8).solve_batch(w.t0, w.tvals, y0, theta)
    yo, so, _ = Oracle(prob, rtol=1e-8, atol=1e-8).solve_forward(w.t0, w.tvals, y0, theta)
    assert (status == 0).all() and (so == 0).all()
    for i in range(B):
        truth = _radau_truth(w, prob, y0[i], theta[i])
        tol = 1e-8 * np.abs(truth) + 1e-8
        err_gpu = np.max(np.abs(y[i] - truth) / tol)
        err_orc = np.max(np.abs(yo[i] - truth) / tol)
        assert err_gpu <= env, (name, i, err_gpu)
        assert err_gpu <= 2.0 * max(err_orc, 5.0), (name, i, err_gpu, err_orc)


def test_robertson_gradients_against_tight_solve():
    """Stiff adjoint: GPU gradients at 1e-8 against the oracle at 1e-11 (forward and backward),
    relative to the oracle@1e-8's own distance from that tight solve."""
    from oracle.oracle import Oracle
    w = examples.workloads()['robertson_adj']
    prob = w.make_problem()
    B = 16
    y0, theta = w.draws(B)
    grads = np.random.default_rng(7).standard_normal((B, len(w.tvals), prob.n_states))
    solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=w.history_capacity)
    _, g, lam, status = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads)
    _, go, lo, so, _ = Oracle(prob, rtol=1e-8, atol=1e-8).solve_adjoint(w.t0, w.tvals, y0, theta, grads)
    _, gt, lt, stt, _ = Oracle(prob, rtol=1e-11, atol=1e-11, rtol_b=1e-12, atol_b=1e-12,
                               rtol_q=1e-12, atol_q=1e-12, mxstep=5000, mxstep_b=5000).solve_adjoint(
        w.t0, w.tvals, y0, theta, grads)
    assert (status == 0).all() and (so == 0).all() and (stt == 0).all()
    # the oracle at 1e-8 is itself 4.5e-7 (gradients) / 2.1e-4 (lamda(t0)) away from the tight
    # solve on these draws: the caps are 20x / 5x that, the relative condition is the check proper
    for a_gpu, a_orc, a_true, cap in ((g, go, gt, 1e-5), (lam, lo, lt, 1e-3)):
        scale = np.abs(a_true).max(axis=0)
        e_gpu = np.max(np.abs(a_gpu - a_true) / scale)
        e_orc = np.max(np.abs(a_orc - a_true) / scale)
        assert e_gpu <= cap, e_gpu
        assert e_gpu <= 2.0 * max(e_orc, 1e-7), (e_gpu, e_orc)


def test_g6_sundials_roberts_example_on_the_gpu():
    """SUNDIALS' own dense Robertson example (cvRoberts_dns: rtol 1e-4, per-state atol, t up to
    4e10; tests/test_oracle.py::test_g6 holds the published statistics and the provenance caveat)
    through `sb_forward`: batch-mean work counters within 10 % of the published CVODE run and
    within 3 % of the oracle's on the same 64 draws, no convergence failure, and -- per draw --
    the oracle's solution to a few tolerance units in the transient (the two step sequences
    decorrelate on this problem: the step count of ONE draw scatters by +-15 %)."""
    from oracle.oracle import Oracle
    from tests.test_oracle import check_roberts_dns_counters, roberts_dns_inputs
    prob = examples.robertson()
    tv, atol, y0, th = roberts_dns_inputs()
    stats = np.zeros((len(y0), 8), dtype=np.int32)
    y, status = Solver(prob, abstol=atol, reltol=1e-4).solve_batch(0.0, tv, y0, th, stats=stats,
                                                                  max_retries=10)
    yo, so, sto = Oracle(prob, rtol=1e-4, atol=atol, mxstep=5000).solve_forward(0.0, tv, y0, th)
    assert (status == 0).all() and (so == 0).all()
    mean = check_roberts_dns_counters(stats)
    mean_o = sto[:, :7].astype(float).mean(axis=0)
    for k in (0, 1, 3, 6):
        assert abs(mean[k] - mean_o[k]) <= 0.03 * mean_o[k], (mean, mean_o)
    units = np.abs(y - yo) / (1e-4 * np.abs(yo) + atol)
    assert units[:, :6].max() <= 10.0, units[:, :6].max()    # transient (observed: 5.7)
    assert units.max() <= 30.0, units.max()                  # tail
