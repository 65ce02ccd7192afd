"""CPU tier: solver options beyond the default path -- Hermite interpolation
of the forward solution (``AdjointSolver(interpolation='hermite')``, reference
solver.py:581-586) -- for the oracle and for the device sources compiled for the host."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from sunode_b200 import examples
from tests.emu.emu import Emulator


def _case(name, B, seed=3):
    w = examples.workloads()[name]
    prob = w.make_problem()
    y0, theta = w.draws(B)
    grads = np.random.default_rng(seed).standard_normal((B, len(w.tvals), prob.n_states))
    return w, prob, y0, theta, grads


def test_oracle_hermite_is_a_valid_adjoint():
    """CV_HERMITE and CV_POLYNOMIAL interpolate the same stored steps: the forward pass is
    bit-identical, the gradients differ at interpolation-error level and both agree with a
    1e-12 solve to the envelope SURVEY.md §8(c) states (1e-5 relative at 1e-8)."""
    w, prob, y0, theta, grads = _case('lv_adj', 16)
    res = {}
    for interp in ('polynomial', 'hermite'):
        o = Oracle(prob, rtol=1e-8, atol=1e-8, interpolation=interp)
        res[interp] = o.solve_adjoint(w.t0, w.tvals, y0, theta, grads)
        assert (res[interp][3] == 0).all()
    np.testing.assert_array_equal(res['hermite'][0], res['polynomial'][0])
    assert not np.array_equal(res['hermite'][1], res['polynomial'][1])
    tight = Oracle(prob, rtol=1e-12, atol=1e-12, rtol_b=1e-12, atol_b=1e-12, rtol_q=1e-12,
                   atol_q=1e-12, mxstep=5000, mxstep_b=5000)
    _, gt, lt, st, _ = tight.solve_adjoint(w.t0, w.tvals, y0, theta, grads)
    assert (st == 0).all()
    for interp in res:
        assert np.max(np.abs(res[interp][1] - gt) / np.abs(gt).max(axis=0)) <= 1e-5
        assert np.max(np.abs(res[interp][2] - lt) / np.abs(lt).max(axis=0)) <= 1e-5


@pytest.mark.parametrize('name,group', [('lv_adj', False), ('seir_adj', False), ('seir_adj', True)])
def test_device_hermite_matches_oracle(name, group, tmp_path):
    """The SB_HERMITE build of the device code (history with y', cubic table entries, unchanged
    backward integrator; one lane per instance and lane groups) takes the oracle's steps."""
    w, prob, y0, theta, grads = _case(name, 12 if group else 32)
    emu = Emulator(prob, str(tmp_path), defines=('SB_HERMITE',), group=group)
    r = emu.adjoint(w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=w.history_capacity,
                    group=group)
    yo, go, lo, so, sto = Oracle(prob, rtol=1e-8, atol=1e-8, interpolation='hermite').solve_adjoint(
        w.t0, w.tvals, y0, theta, grads)
    assert (r['status'] == 0).all() and (so == 0).all()
    assert np.max(np.abs(r['y'] - yo) / (1e-8 * np.abs(yo) + 1e-8)) <= 1e-3
    assert np.max(np.abs(r['grad'] - go) / np.abs(go).max(axis=0)) <= 1e-9
    assert np.max(np.abs(r['lamda'] - lo) / np.abs(lo).max(axis=0)) <= 1e-9
    assert (r['stats'][:, 0] == sto[:, 7]).mean() >= 0.9
    # and the polynomial oracle gives a (slightly) different answer: the option is not a no-op
    gp = Oracle(prob, rtol=1e-8, atol=1e-8).solve_adjoint(w.t0, w.tvals, y0, theta, grads)[1]
    assert np.max(np.abs(r['grad'] - gp) / np.abs(gp).max(axis=0)) > 1e-12


def test_hermite_table_entries(tmp_path):
    """Every table entry of the SB_HERMITE build is the cubic through (y, y') at both ends of its
    step: checked by evaluating the entry the way the backward kernels do (Newton form, nodes
    T[i], scaled by 1/delt) at the ends and by a finite difference of it for the slopes."""
    w, prob, y0, theta, _ = _case('lv_adj', 2)
    emu = Emulator(prob, str(tmp_path), defines=('SB_HERMITE',))
    r = emu.adjoint(w.t0, w.tvals, y0, theta, np.ones((50, 2)), 1e-8, 1e-8, hist_cap=512)
    ns = 2

    def evaluate(e, t):
        order, inv = int(e[2]), e[3]
        y = e[10:10 + ns].copy()
        c = 1.0
        for i in range(order):
            c *= (t - e[4 + i]) * inv
            y += c * e[10 + ns * (i + 1):10 + ns * (i + 2)]
        return y

    for b in range(2):
        n = r['fwd']['hist_n'][b]
        hist, tab = r['fwd']['hist'][b, :n], r['tab'][b]
        assert hist.shape[1] == 2 * ns + 2
        for idx in range(1, n):
            e = tab[idx]
            t0, t1 = hist[idx - 1, 0], hist[idx, 0]
            assert e[0] == t0 and e[1] == t1 and e[2] == 3.0
            d = t1 - t0
            np.testing.assert_allclose(evaluate(e, t1), hist[idx, 2:2 + ns], rtol=0, atol=0)
            np.testing.assert_allclose(evaluate(e, t0), hist[idx - 1, 2:2 + ns], rtol=1e-13)
            eps = 1e-6 * d
            s1 = (evaluate(e, t1 + eps) - evaluate(e, t1 - eps)) / (2 * eps)
            s0 = (evaluate(e, t0 + eps) - evaluate(e, t0 - eps)) / (2 * eps)
            np.testing.assert_allclose(s1, hist[idx, 2 + ns:], rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(s0, hist[idx - 1, 2 + ns:], rtol=1e-6, atol=1e-9)
