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
    w, prob, y0, theta, grads = _case(name, 4 if group else 32)
    emu = Emulator(prob, str(tmp_path), defines=('SB_HERMITE',), group=group)
    r = emu.adjoint(w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=w.history_capacity,
                    group=group)
    yo, go, lo, so, sto = Oracle(prob, rtol=1e-8, atol=1e-8, interpolation='hermite').solve_adjoint(
        w.t0, w.tvals, y0, theta, grads)
    assert (r['status'] == 0).all() and (so == 0).all()
    assert np.max(np.abs(r['y'] - yo) / (1e-8 * np.abs(yo) + 1e-8)) <= 1e-3
    assert np.max(np.abs(r['grad'] - go) / np.abs(go).max(axis=0)) <= 1e-9
    assert np.max(np.abs(r['lamda'] - lo) / np.abs(lo).max(axis=0)) <= 1e-9
    # same step sequence up to the odd rounding-level flip of a controller decision
    assert np.max(np.abs(r['stats'][:, 0] - sto[:, 7]) / sto[:, 7]) <= 0.02
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


# ---------------------------------------------------------------------------------- constraints
def chase_problem():
    """Two states that test ``constraints``: ``a`` relaxes towards 0.6 + cos(t), which dips below
    zero (a constraint on it cannot be met: CV_CONSTR_FAIL / CV_CONV_FAILURE after ten shrunken
    steps), ``b`` decays towards zero, where at loose tolerances the unconstrained integrator
    undershoots (a constraint on it is met by projection or by a smaller step)."""
    import sympy as sy
    from sunode_b200 import SympyProblem

    def rhs(t, y, p):
        return {'a': -p.k * (y.a - (0.6 + sy.cos(t))),
                'b': -p.m * y.b * y.b - p.k * y.b * y.a}
    return SympyProblem(params={'k': (), 'm': ()}, states={'a': (), 'b': ()}, rhs_sympy=rhs,
                        derivative_params=[('k',), ('m',)])


def chase_inputs(B=16):
    rng = np.random.default_rng(0)
    theta = np.array([20.0, 3.0]) * np.exp(0.3 * rng.standard_normal((B, 2)))
    return np.array([1.6, 1.0]), theta, np.linspace(0.1, 6, 40)


def constraint_define(cons):
    return 'SB_CONSTRAINTS=' + ','.join('%.1f' % c for c in cons)


def test_oracle_constraints_change_the_solution():
    prob = chase_problem()
    y0, theta, tv = chase_inputs()
    free = Oracle(prob, rtol=1e-4, atol=1e-7).solve_forward(0.0, tv, y0, theta)
    assert (free[1] == 0).all() and free[0][..., 1].min() < -1e-3      # undershoots without
    con = Oracle(prob, rtol=1e-4, atol=1e-7, constraints=[0.0, 1.0]).solve_forward(0.0, tv, y0, theta)
    assert (con[1] == 0).all()
    assert con[0][..., 1].min() > -1e-6                  # dense output between non-negative steps
    # a constraint the true solution violates cannot be met
    bad = Oracle(prob, rtol=1e-4, atol=1e-7, constraints=[1.0, 0.0]).solve_forward(0.0, tv, y0, theta)
    assert np.isin(bad[1], (-15, -4)).all() and (bad[1] == -15).any()
    assert np.isnan(bad[0]).all()
    # cvInitialSetup: y0 must satisfy the constraints
    st0 = Oracle(prob, rtol=1e-4, atol=1e-7, constraints=[0.0, 2.0]).solve_forward(
        0.0, tv, np.array([1.6, 0.0]), theta[:2])[1]
    assert (st0 == -22).all()


@pytest.mark.parametrize('cons', [[0.0, 1.0], [0.0, 2.0], [2.0, 1.0]])
def test_device_constraints_match_oracle(cons, tmp_path):
    """The SB_CONSTRAINTS build of the forward integrator (Bdf::check_constraints) against the
    oracle's cvCheckConstraints: same outcomes per draw (including which draws give up with
    CV_CONSTR_FAIL and which with CV_CONV_FAILURE), same step / failure counters except where a
    rounding-level difference flips a sign test."""
    prob = chase_problem()
    y0, theta, tv = chase_inputs()
    emu = Emulator(prob, str(tmp_path), defines=(constraint_define(cons),))
    r = emu.forward(0.0, tv, y0, theta, 1e-4, 1e-7)
    yo, so, sto = Oracle(prob, rtol=1e-4, atol=1e-7, constraints=cons).solve_forward(0.0, tv, y0, theta)
    np.testing.assert_array_equal(r['status'], so)
    assert (r['stats'][:, 0] == sto[:, 0]).mean() >= 0.8
    assert (r['stats'][:, 5] == sto[:, 5]).mean() >= 0.8
    ok = so == 0
    if ok.any():
        assert np.max(np.abs(r['y'][ok] - yo[ok]) / (1e-4 * np.abs(yo[ok]) + 1e-7)) <= 100.0
    assert np.isnan(r['y'][~ok]).all()
    bad0 = emu.forward(0.0, tv, np.array([1.6, -0.5]), theta[:2], 1e-4, 1e-7)
    assert (bad0['status'] == -22).all()


def test_constraints_leave_the_backward_pass_alone(tmp_path):
    """AdjointSolver(constraints=...) constrains the forward ODE only (reference
    solver.py:566-572 sets them on ``self._ode``): with flags that never bind the constrained
    build reproduces the unconstrained adjoint bit for bit -- one lane per instance and lane
    groups (3 draws: the group emulation runs its lanes as threads)."""
    for name, group in (('lv_adj', False), ('seir_adj', True)):
        w, prob, y0, theta, grads = _case(name, 3)
        cons = [1.0] * prob.n_states
        ref = Emulator(prob, str(tmp_path), group=group).adjoint(
            w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=w.history_capacity, group=group)
        con = Emulator(prob, str(tmp_path), defines=(constraint_define(cons),), group=group).adjoint(
            w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=w.history_capacity, group=group)
        assert (con['status'] == 0).all()
        np.testing.assert_array_equal(con['y'], ref['y'])
        np.testing.assert_array_equal(con['grad'], ref['grad'])
        np.testing.assert_array_equal(con['lamda'], ref['lamda'])


# ---------------------------------------------------------------------------------- scaling factors
def test_sens_scaling_factors(tmp_path):
    """``Solver(scaling_factors=pbar)`` (reference solver.py:381-389: CVodeSetSensParams +
    CVodeSensEEtolerances): sensitivity block k is controlled with atol / |pbar_k|.  Oracle:
    pbar = 1 is the default run bit for bit, a large pbar tightens the sensitivity error control
    (more steps); emulated device code = oracle."""
    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    y0, theta = w.draws(8)
    s0 = np.zeros((2, 2))
    base = Oracle(prob, rtol=1e-6, atol=1e-6).solve_forward_sens(w.t0, w.tvals, y0, theta, s0)
    ones = Oracle(prob, rtol=1e-6, atol=1e-6, scaling_factors=np.ones(2)).solve_forward_sens(
        w.t0, w.tvals, y0, theta, s0)
    np.testing.assert_array_equal(base[1], ones[1])
    pbar = np.array([1e4, -1e3])
    scaled = Oracle(prob, rtol=1e-6, atol=1e-6, scaling_factors=pbar).solve_forward_sens(
        w.t0, w.tvals, y0, theta, s0)
    assert (scaled[2] == 0).all() and (scaled[3][:, 0] >= base[3][:, 0]).all()
    assert not np.array_equal(scaled[1], base[1])
    from tests.emu.emu import Emulator
    r = Emulator(prob, str(tmp_path)).forward_sens(w.t0, w.tvals, y0, theta, s0, 1e-6, 1e-6, pbar=pbar)
    assert (r['status'] == 0).all()
    np.testing.assert_array_equal(r['stats'][:, 0], scaled[3][:, 0])
    np.testing.assert_allclose(r['sens'], scaled[1], rtol=1e-7, atol=1e-9 * np.abs(scaled[1]).max())


def test_constraint_flags_to_build_option():
    from sunode_b200.solver import _constraint_defines
    assert _constraint_defines(None, 3) == (None, ())
    c, d = _constraint_defines(0.0, 2)
    assert d == () and c.shape == (2,)                           # all-zero flags: the default build
    c, d = _constraint_defines(np.array([1, 0, -2]), 3)
    assert d == ('SB_CONSTRAINTS=1.0,0.0,-2.0',)
    with pytest.raises(ValueError, match='CV_ILL_INPUT'):
        _constraint_defines([3.0, 0.0], 2)


# ---------------------------------------------------------------------------------- restart-free backward
@pytest.mark.parametrize('name', ['lv_adj', 'robertson_adj'])
def test_fundamental_matrix_backward_pass(name, tmp_path):
    """SURVEY.md 8(f) #3, ``AdjointSolver(backward='fundamental')`` (csrc/sb_fund.cuh): the
    fundamental matrix of the adjoint equation integrated without restarts, jumps as dense solves.
    Same gradients as the reference's restart-per-output-time schedule to the tolerances (both
    are equally far from a 1e-12 solve); on the smooth problem a sixth of the backward steps and
    no re-basing, on the stiff one the conditioning guard re-bases at most output times."""
    w, prob, y0, theta, grads = _case(name, 16)
    emu = Emulator(prob, str(tmp_path), defines=('SB_FUND',))
    r = emu.adjoint(w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=w.history_capacity, fund=True)
    ref = Oracle(prob, rtol=1e-8, atol=1e-8).solve_adjoint(w.t0, w.tvals, y0, theta, grads)
    tight = Oracle(prob, rtol=1e-12, atol=1e-12, rtol_b=1e-12, atol_b=1e-12, rtol_q=1e-12,
                   atol_q=1e-12, mxstep=20000, mxstep_b=20000).solve_adjoint(w.t0, w.tvals, y0, theta, grads)
    assert (r['status'] == 0).all() and (ref[3] == 0).all() and (tight[3] == 0).all()
    gs, ls = np.abs(tight[1]).max(axis=0), np.abs(tight[2]).max(axis=0)
    err_fund = np.max(np.abs(r['grad'] - tight[1]) / gs)
    err_ref = np.max(np.abs(ref[1] - tight[1]) / gs)
    assert err_fund <= 1e-5 and err_fund <= 3 * err_ref + 1e-8       # SURVEY 8(c) envelope at 1e-8
    # (lamda(t0) of the stiff problem is the less accurate output of either schedule)
    assert np.max(np.abs(r['lamda'] - tight[2]) / ls) <= max(1e-5, 3 * np.max(np.abs(ref[2] - tight[2]) / ls))
    steps, rebases, steps_ref = r['stats'][:, 0], r['stats'][:, 7], ref[4][:, 7]
    if name == 'lv_adj':
        assert np.max(np.abs(r['grad'] - ref[1]) / gs) <= 1e-7
        assert (rebases == 0).all() and (steps < 0.3 * steps_ref).all()
    else:
        assert (rebases > 10).all() and (steps < 1.5 * steps_ref).all()


def test_fundamental_matrix_backward_edge_cases(tmp_path):
    """Output times that include t0, a last output time before t_start's interval is empty, a
    single output time, all-t0 outputs: the restart-free pass against the reference schedule."""
    w, prob, y0, theta, _ = _case('lv_adj', 4)
    emu = Emulator(prob, str(tmp_path), defines=('SB_FUND',))
    rng = np.random.default_rng(9)
    for tv in (np.linspace(0, 10, 7), np.linspace(0.5, 9, 5), np.array([3.0]), np.array([0.0]),
               np.array([0.0, 0.0, 2.0, 2.0, 5.0])):
        g = rng.standard_normal((4, len(tv), 2))
        r = emu.adjoint(w.t0, tv, y0, theta, g, 1e-8, 1e-8, hist_cap=512, fund=True)
        ref = Oracle(prob, rtol=1e-8, atol=1e-8).solve_adjoint(w.t0, tv, y0, theta, g)
        assert (r['status'] == 0).all() and (ref[3] == 0).all(), tv
        scale = np.abs(ref[1]).max() + 1e-300
        np.testing.assert_allclose(r['grad'], ref[1], rtol=0, atol=2e-7 * scale + 1e-14, err_msg=str(tv))
        np.testing.assert_allclose(r['lamda'], ref[2], rtol=0, atol=2e-7 * np.abs(ref[2]).max() + 1e-14,
                                   err_msg=str(tv))


def test_build_options_combine(tmp_path):
    """Hermite tables + inactive constraints + the restart-free backward pass in one build: the
    options are independent of each other (history layout / forward integrator / backward
    driver), the result is the Hermite oracle's to the restart-free pass's tolerance."""
    w, prob, y0, theta, grads = _case('lv_adj', 8)
    emu = Emulator(prob, str(tmp_path), defines=('SB_HERMITE', 'SB_FUND', constraint_define([1.0, 1.0])))
    r = emu.adjoint(w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=512, fund=True)
    ref = Oracle(prob, rtol=1e-8, atol=1e-8, interpolation='hermite', constraints=[1.0, 1.0]).solve_adjoint(
        w.t0, w.tvals, y0, theta, grads)
    assert (r['status'] == 0).all() and (ref[3] == 0).all()
    np.testing.assert_array_equal(r['fwd']['stats'][:, 0], ref[4][:, 0])
    assert np.max(np.abs(r['grad'] - ref[1]) / np.abs(ref[1]).max(axis=0)) <= 1e-7
    assert np.max(np.abs(r['lamda'] - ref[2]) / np.abs(ref[2]).max(axis=0)) <= 1e-7
    assert (r['stats'][:, 0] < 0.3 * ref[4][:, 7]).all()
