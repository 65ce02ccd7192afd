"""GPU tier: the framework wrappers (PyTensor Ops through ``perform``, torch.autograd) and the
README-style configuration pokes, against the oracle / closed forms."""
import numpy as np
import pytest

from sunode_b200 import SympyProblem, _cvodes, examples
from sunode_b200.solver import AdjointSolver

pytestmark = pytest.mark.gpu


def _oracle(problem, **kw):
    from oracle.oracle import Oracle
    return Oracle(problem, **kw)


def test_pytensor_ops_perform():
    """The reference's Ops, driven the way pytensor's VM drives them (``perform``)."""
    from sunode_b200.wrappers import as_pytensor as ap
    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8)
    theta = np.array([0.1, 0.2, 0.3, 0.4])
    y0 = np.array([1.0, 0.1])
    params, fixed = theta[:2].copy(), theta[2:].copy()      # derivative subset = (alpha, beta)
    out = [[None]]
    ap.SolveODEAdjoint(solver).perform(None, [y0, params, fixed, np.float64(0.0), w.tvals], out)
    y = out[0][0]
    yo, go, lo, so, _ = _oracle(prob, rtol=1e-8, atol=1e-8).solve_adjoint(
        0.0, w.tvals, y0, theta, np.ones((50, 2)))
    tol = 1e-8 * np.abs(yo[0]) + 1e-8
    assert np.max(np.abs(y - yo[0]) / tol) <= 1.0
    out = [[None], [None]]
    ap.SolveODEAdjointBackward(solver).perform(
        None, [y0, params, fixed, np.ones((50, 2)), np.float64(0.0), w.tvals], out)
    np.testing.assert_allclose(out[0][0], lo[0], rtol=1e-7)
    np.testing.assert_allclose(out[1][0], go[0], rtol=1e-7)
    out = [[None]]
    ap.EvalRhs(solver).perform(None, [params, fixed, y, w.tvals], out)
    a, b, c, d = theta
    np.testing.assert_allclose(out[0][0][:, 0], a * y[:, 0] - b * y[:, 0] * y[:, 1], rtol=1e-13)
    # batched Ops
    B = 64
    y0b, thb = w.draws(B)
    out = [[None]]
    ap.SolveODEAdjointBatch(solver).perform(None, [y0b, thb[:, :2], thb[0, 2:], np.float64(0.0), w.tvals], out)
    thb_eff = thb.copy()
    thb_eff[:, 2:] = thb[0, 2:]
    yob, gob, lob, _, _ = _oracle(prob, rtol=1e-8, atol=1e-8).solve_adjoint(
        0.0, w.tvals, y0b, thb_eff, np.ones((50, 2)))
    assert np.max(np.abs(out[0][0] - yob) / (1e-8 * np.abs(yob) + 1e-8)) <= 1.0
    out = [[None], [None]]
    ap.SolveODEAdjointBackwardBatch(solver).perform(
        None, [y0b, thb[:, :2], thb[0, 2:], np.ones((B, 50, 2)), np.float64(0.0), w.tvals], out)
    np.testing.assert_allclose(out[1][0], gob, rtol=1e-6)
    np.testing.assert_allclose(out[0][0], lob, rtol=1e-6)
    # a failing draw gives NaN, not an exception (as_pytensor.py:287-290)
    out = [[None]]
    ap.SolveODEAdjoint(solver).perform(
        None, [np.array([np.nan, 0.1]), params, fixed, np.float64(0.0), w.tvals], out)
    assert np.isnan(out[0][0]).all()


def test_torch_autograd_matches_oracle_gradients():
    torch = pytest.importorskip('torch')
    from sunode_b200.wrappers import as_torch
    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8)
    B = 128
    y0, theta = w.draws(B)
    dev = torch.device('cuda:0')
    y0_t = torch.tensor(y0, device=dev, requires_grad=True)
    pd_t = torch.tensor(theta[:, :2], device=dev, requires_grad=True)
    pf_t = torch.tensor(theta[:, 2:], device=dev)
    weights = torch.tensor(np.random.default_rng(1).standard_normal((B, 50, 2)), device=dev)
    y, status = as_torch.solve_ivp(solver, 0.0, w.tvals, y0_t, pd_t, pf_t)
    assert (status == 0).all()
    loss = (weights * y).sum()
    loss.backward()
    yo, go, lo, _, _ = _oracle(prob, rtol=1e-8, atol=1e-8).solve_adjoint(
        0.0, w.tvals, y0, theta, weights.cpu().numpy())
    np.testing.assert_allclose(pd_t.grad.cpu().numpy(), go, rtol=1e-6, atol=1e-9 * np.abs(go).max())
    np.testing.assert_allclose(y0_t.grad.cpu().numpy(), -lo, rtol=1e-6, atol=1e-9 * np.abs(lo).max())


def test_readme_style_pokes_change_the_backward_tolerances():
    """README.md:243-249 of the reference, verbatim calls."""
    def rhs(t, y, p):
        return {'x': -p.k * y.x ** 2}

    prob = SympyProblem({'k': ()}, {'x': ()}, rhs, [('k',)])
    solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8)
    lib = _cvodes.lib
    tvals = np.linspace(0.1, 3, 30)
    y0, k = np.ones((1, 1)), np.full((1, 1), 0.7)
    g = np.ones((30, 1))
    sb = np.zeros((1, 8), np.int32)
    solver.solve_adjoint_batch(0.0, tvals, y0, k, g, stats_bwd=sb)
    steps_tight = sb[0, 0]
    lib.CVodeSStolerancesB(solver._ode, solver._odeB, 1e-5, 1e-5)
    lib.CVodeQuadSStolerancesB(solver._ode, solver._odeB, 1e-5, 1e-5)
    lib.CVodeSetMaxNumSteps(solver._ode, 5000)
    lib.CVodeSetMaxNumStepsB(solver._ode, solver._odeB, 5000)
    _, grad, lam, st = solver.solve_adjoint_batch(0.0, tvals, y0, k, g, stats_bwd=sb)
    assert st[0] == 0 and sb[0, 0] < steps_tight
    # x(t) = 1 / (1 + k t): dL/dk = sum_i -t_i / (1 + k t_i)^2
    np.testing.assert_allclose(grad[0, 0], np.sum(-tvals / (1 + 0.7 * tvals) ** 2), rtol=1e-3)
