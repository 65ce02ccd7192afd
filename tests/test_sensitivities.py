"""Forward sensitivity analysis (reference solver.py:360-392, 483-527):
CPU tier pins the oracle (closed form, finite differences, consistency with the adjoint) and
checks the emulated device code against it; the GPU tier goes through the C ABI."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from sunode_b200 import SympyProblem, examples


def _smoke_problem():
    return SympyProblem({'a': {'b': ()}}, {'x': ()}, lambda t, y, p: {'x': y.x + p.a.b},
                        [('a', 'b')])


def test_oracle_closed_form():
    """x' = x + b, x(0) = 1: dx/db = e^t - 1 (the reference's smoke problem, test_solve.py:81-117)."""
    prob = _smoke_problem()
    t = np.linspace(0, 1)
    y, s, st, _ = Oracle(prob).solve_forward_sens(0.0, t, [1.0], [0.2], np.zeros((1, 1)))
    assert st[0] == 0
    np.testing.assert_allclose(y[0, :, 0], 1.2 * np.exp(t) - 0.2, rtol=1e-8)
    np.testing.assert_allclose(s[0, 1:, 0, 0], (np.exp(t) - 1)[1:], rtol=1e-7)
    # sensitivity w.r.t. the initial value: seed the unit vector instead (as_pytensor.py:211-230)
    y, s, st, _ = Oracle(prob).solve_forward_sens(0.0, t, [1.0], [0.2], np.ones((1, 1)))
    np.testing.assert_allclose(s[0, :, 0, 0], 2 * np.exp(t) - 1, rtol=1e-7)   # e^t + (e^t - 1)


@pytest.mark.parametrize('name', ['lv_adj', 'robertson_adj'])
def test_oracle_sensitivities_vs_finite_differences_and_adjoint(name):
    w = examples.workloads()[name]
    prob = w.make_problem()
    y0, theta = w.draws(1)
    n_s, n_d = prob.n_states, prob.n_params
    orc = Oracle(prob, rtol=1e-8, atol=1e-8)
    y, s, st, _ = orc.solve_forward_sens(w.t0, w.tvals, y0, theta, np.zeros((n_d, n_s)))
    assert st[0] == 0
    tight = Oracle(prob, rtol=1e-12, atol=1e-14, mxstep=200000)
    for k, idx in enumerate(prob.generated.deriv_index):
        h = 1e-5 * theta[0, idx]
        tp, tm = theta.copy(), theta.copy()
        tp[0, idx] += h
        tm[0, idx] -= h
        fd = (tight.solve_forward(w.t0, w.tvals, y0, tp)[0][0]
              - tight.solve_forward(w.t0, w.tvals, y0, tm)[0][0]) / (2 * h)
        scale = np.abs(fd).max()
        assert np.max(np.abs(s[0, :, k, :] - fd)) <= 3e-5 * scale
    g = np.random.default_rng(0).standard_normal((len(w.tvals), n_s))
    _, grad, _, st2, _ = orc.solve_adjoint(w.t0, w.tvals, y0, theta, g)
    np.testing.assert_allclose(np.einsum('ti,tki->k', g, s[0]), grad[0], rtol=2e-4)


def test_emulated_device_code_matches_oracle(tmp_path):
    from tests.emu.emu import Emulator
    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    y0, theta = w.draws(32)
    sens0 = np.random.default_rng(2).standard_normal((32, 2, 2))
    r = Emulator(prob, str(tmp_path)).forward_sens(w.t0, w.tvals, y0, theta, sens0, 1e-8, 1e-8)
    y, s, st, stats = Oracle(prob, rtol=1e-8, atol=1e-8).solve_forward_sens(
        w.t0, w.tvals, y0, theta, sens0)
    assert (r['status'] == 0).all() and (st == 0).all()
    assert np.max(np.abs(r['y'] - y) / (1e-8 * np.abs(y) + 1e-8)) <= 1e-3
    assert np.max(np.abs(r['sens'] - s) / (1e-8 * np.abs(s) + 1e-8)) <= 1e-3
    assert (r['stats'][:, 0] == stats[:, 0]).mean() >= 0.9


def test_initial_value_pseudo_parameters():
    from sunode_b200.wrappers import as_pytensor as ap
    prob = SympyProblem({'k': (), '__initial_values': {'x': (), 'v': 2}}, {'x': (), 'v': 2},
                        lambda t, y, p: {'x': -p.k * y.x, 'v': [y.v[1], -y.v[0]]},
                        [('k',), ('__initial_values', 'v')])
    np.testing.assert_array_equal(ap.initial_sensitivities(prob),
                                  [[0, 0, 0], [0, 1, 0], [0, 0, 1]])


@pytest.mark.gpu
def test_gpu_forward_sensitivities_match_oracle():
    from sunode_b200.solver import Solver
    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    B = 256
    y0, theta = w.draws(B)
    sens0 = np.zeros((2, 2))
    solver = Solver(prob, abstol=1e-8, reltol=1e-8, sens_mode='simultaneous')
    y, s, st = solver.solve_sens_batch(w.t0, w.tvals, y0, theta, sens0)
    yo, so, sto, _ = Oracle(prob, rtol=1e-8, atol=1e-8).solve_forward_sens(
        w.t0, w.tvals, y0, theta, sens0)
    assert (st == 0).all() and (sto == 0).all()
    assert np.max(np.abs(y - yo) / (1e-8 * np.abs(yo) + 1e-8)) <= 1.0
    assert np.max(np.abs(s - so) / (1e-8 * np.abs(so) + 1e-8)) <= 1.0


@pytest.mark.gpu
def test_gpu_reference_shaped_sensitivity_api():
    """check_call_solve of the reference (test_solve.py:81-117) for both sens modes, with the
    closed form attached; and the SolveODE op's numeric body."""
    from sunode_b200.solver import Solver
    from sunode_b200.wrappers import as_pytensor as ap
    prob = _smoke_problem()
    t = np.linspace(0, 1)
    for mode in ('simultaneous', 'staggered'):
        solver = Solver(prob, sens_mode=mode)
        solver.set_params_dict({'a': {'b': 0.2}})
        y_out, sens_out = solver.make_output_buffers(t)
        assert sens_out.shape == (50, 1, 1)
        solver.solve(0, t, np.ones(1), y_out, sens0=np.zeros((1, 1)), sens_out=sens_out)
        np.testing.assert_allclose(y_out[:, 0], 1.2 * np.exp(t) - 0.2, rtol=1e-7)
        np.testing.assert_allclose(sens_out[1:, 0, 0], (np.exp(t) - 1)[1:], rtol=1e-6)
        with pytest.raises(ValueError):
            solver.solve(0, t, np.ones(1), y_out)
    with pytest.raises(ValueError):
        Solver(prob, sens_mode='staggered1')
    out = [[None], [None]]
    ap.SolveODE(solver).perform(None, [np.ones(1), np.array([0.2]), np.zeros(0), np.float64(0.0), t], out)
    np.testing.assert_allclose(out[1][0][1:, 0, 0], (np.exp(t) - 1)[1:], rtol=1e-6)


@pytest.mark.gpu
def test_gpu_forward_sensitivities_in_lane_groups_match_oracle():
    """The 8-state SEIR problem: y and the six sensitivity vectors dy/dp_k in lane groups of
    4 lanes x 2 components per block (csrc/sb_group.cuh, forward_instance_group) against the
    oracle; and the grouped build against the one-lane-per-instance build of the same kernel."""
    import os
    from sunode_b200.solver import Solver
    w = examples.workloads()['seir_adj']
    prob = w.make_problem()
    B = 70                                           # not a multiple of 8 groups per warp
    y0, theta = w.draws(B)
    sens0 = np.zeros((prob.n_params, prob.n_states))
    y, s, st = Solver(prob, abstol=1e-8, reltol=1e-8, sens_mode='simultaneous').solve_sens_batch(
        w.t0, w.tvals, y0, theta, sens0)
    yo, so, sto, _ = Oracle(prob, rtol=1e-8, atol=1e-8).solve_forward_sens(w.t0, w.tvals, y0, theta, sens0)
    assert (st == 0).all() and (sto == 0).all()
    assert np.max(np.abs(y - yo) / (1e-8 * np.abs(yo) + 1e-8)) <= 1.0
    assert np.max(np.abs(s - so) / (1e-8 * np.abs(so) + 1e-8)) <= 1.0
    saved = os.environ.get('SUNODE_B200_DEFINES')
    os.environ['SUNODE_B200_DEFINES'] = 'SB_NO_FWD_GROUP'
    try:
        y1, s1, st1 = Solver(prob, abstol=1e-8, reltol=1e-8, sens_mode='simultaneous').solve_sens_batch(
            w.t0, w.tvals, y0, theta, sens0)
    finally:
        if saved is None:
            del os.environ['SUNODE_B200_DEFINES']
        else:
            os.environ['SUNODE_B200_DEFINES'] = saved
    assert (st1 == 0).all()
    assert np.max(np.abs(y - y1) / (1e-8 * np.abs(y1) + 1e-8)) <= 1e-2
    assert np.max(np.abs(s - s1) / (1e-8 * np.abs(s1) + 1e-8)) <= 1e-2
