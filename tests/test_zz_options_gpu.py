"""GPU tier (``-m gpu``), solver options beyond the default path; kept in a
file of its own that sorts last so that the parity tests of the headline path run first."""
import numpy as np
import pytest

from sunode_b200 import examples
from sunode_b200.solver import AdjointSolver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['lv_adj', 'seir_adj'])
def test_hermite_interpolation_matches_oracle(name):
    """``AdjointSolver(interpolation='hermite')`` (reference solver.py:581-586): the SB_HERMITE
    build of the kernels against the oracle's CV_HERMITE, same envelope as the polynomial path
    (trajectories 1 tolerance unit, gradients 1e-7 relative)."""
    from oracle.oracle import Oracle
    w = examples.workloads()[name]
    prob = w.make_problem()
    B = 256
    y0, theta = w.draws(B)
    grads = np.random.default_rng(5).standard_normal((B, len(w.tvals), prob.n_states))
    solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, interpolation='hermite',
                           history_capacity=w.history_capacity)
    y, g, lam, status = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads)
    yo, go, lo, so, _ = Oracle(prob, rtol=1e-8, atol=1e-8, interpolation='hermite').solve_adjoint(
        w.t0, w.tvals, y0, theta, grads)
    assert (status == 0).all() and (so == 0).all()
    assert np.max(np.abs(y - yo) / (1e-8 * np.abs(yo) + 1e-8)) <= 1.0
    assert np.max(np.abs(g - go) / np.abs(go).max(axis=0)) <= 1e-7
    assert np.max(np.abs(lam - lo) / np.abs(lo).max(axis=0)) <= 1e-7
    # not the polynomial answer
    plain = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=w.history_capacity)
    gp = plain.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads)[1]
    assert np.max(np.abs(g - gp) / np.abs(gp).max(axis=0)) > 1e-12
