"""CPU tier: the DEVICE integrator sources (csrc/sb_bdf.cuh, sb_kernels.cuh + the generated
__device__ functions) compiled for the host by tests/emu and compared with the oracle.  This
checks the kernel logic where no GPU is available; the `-m gpu` tier runs the same comparison on
the real kernels through the C ABI."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from sunode_b200 import examples
from tests.emu.emu import Emulator


@pytest.mark.parametrize('name,env', [('lv_adj', 1e-3), ('seir_adj', 1e-3), ('robertson_adj', 1000.0)])
def test_device_logic_matches_oracle(name, env, tmp_path):
    w = examples.workloads()[name]
    prob = w.make_problem()
    B = 48
    y0, theta = w.draws(B)
    grads = np.random.default_rng(11).standard_normal((B, len(w.tvals), prob.n_states))
    emu = Emulator(prob, str(tmp_path))
    r = emu.adjoint(w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=w.history_capacity)
    yo, go, lo, so, sto = Oracle(prob, rtol=1e-8, atol=1e-8).solve_adjoint(
        w.t0, w.tvals, y0, theta, grads)
    assert (r['status'] == 0).all() and (so == 0).all()
    tol = 1e-8 * np.abs(yo) + 1e-8
    assert np.max(np.abs(r['y'] - yo) / tol) <= env
    gtol = 1e-9 if env < 1 else 1e-5
    assert np.max(np.abs(r['grad'] - go) / np.abs(go).max(axis=0)) <= gtol
    assert np.max(np.abs(r['lamda'] - lo) / np.abs(lo).max(axis=0)) <= max(gtol, 1e-4 if env > 1 else 0)
    if env < 1:
        # same step sequence: identical counters, except where a rounding-level difference (the
        # device code uses FMAs and its own k-th root) flips a controller decision
        np.testing.assert_array_equal(r['fwd']['stats'][:, 0], sto[:, 0])
        same = r['stats'][:, 0] == sto[:, 7]
        assert same.mean() >= 0.9
        assert np.max(np.abs(r['stats'][:, 0] - sto[:, 7]) / sto[:, 7]) <= 0.02


def test_history_and_tables_reproduce_forward_solution(tmp_path):
    """The interpolation tables built from the stored steps (sb_tables) evaluate to the stored
    points at the step times, and the stored history equals the oracle's data points."""
    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    y0, theta = w.draws(1)
    emu = Emulator(prob, str(tmp_path))
    r = emu.adjoint(w.t0, w.tvals, y0, theta, np.ones((50, 2)), 1e-8, 1e-8, hist_cap=512)
    n = r['fwd']['hist_n'][0]
    hist = r['fwd']['hist'][0, :n]
    st, _, ht, ho, hy = Oracle(prob, rtol=1e-8, atol=1e-8).forward_history(
        w.t0, w.tvals, y0, theta)
    assert st == 0 and len(ht) == n
    # same step sequence; the step sizes agree to rounding-error propagation (the device code
    # uses explicit FMAs, the oracle plain arithmetic)
    np.testing.assert_allclose(hist[:, 0], ht, rtol=1e-8)
    np.testing.assert_array_equal(hist[1:, 1].astype(int), ho[1:])
    np.testing.assert_allclose(hist[:, 2:], hy, rtol=1e-8)
    tab = r['tab'][0]
    for idx in range(1, n):
        e = tab[idx]
        assert e[0] == hist[idx - 1, 0] and e[1] == hist[idx, 0]
        np.testing.assert_array_equal(e[10:12], hist[idx, 2:])               # Y[0] = right end point


def test_tolerance_mxstep_and_capacity_limits(tmp_path):
    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    y0, theta = w.draws(4)
    emu = Emulator(prob, str(tmp_path))
    r = emu.forward(w.t0, w.tvals, y0, theta, 1e-8, 1e-8, max_steps=1)       # 1 step per tval
    assert (r['status'] == -1).all() and np.isnan(r['y']).all()              # CV_TOO_MUCH_WORK
    r = emu.forward(w.t0, w.tvals, y0, theta, 1e-8, 1e-8, hist_cap=16, max_steps=2500)
    assert (r['status'] == -1).all()                                         # history overflow
    r = emu.forward(w.t0, w.tvals, y0, theta, 0.0, 0.0)                      # ewt undefined
    assert (r['status'] == -22).all()                                        # CV_ILL_INPUT


def test_edge_cases_match_oracle(tmp_path):
    """No parameters at all, no derivative parameters, a single output time equal to t0 (no
    interval to integrate: only the jump is applied, solver.py:750-776), t0 among the outputs."""
    from sunode_b200 import SympyProblem
    prob = SympyProblem({}, {'x': ()}, lambda t, y, p: {'x': -y.x}, [])
    emu = Emulator(prob, str(tmp_path / 'a'))
    t = np.linspace(0.1, 1, 5)
    r = emu.adjoint(0.0, t, np.ones((3, 1)), np.zeros((3, 0)), np.ones((5, 1)), 1e-8, 1e-8, hist_cap=128)
    assert (r['status'] == 0).all() and r['grad'].shape == (3, 0)
    np.testing.assert_allclose(-r['lamda'][:, 0], np.sum(np.exp(-t)), rtol=1e-6)

    prob = SympyProblem({'k': ()}, {'x': ()}, lambda t, y, p: {'x': -p.k * y.x}, [])
    emu = Emulator(prob, str(tmp_path / 'b'))
    orc = Oracle(prob, rtol=1e-8, atol=1e-8)
    for tv in (np.array([0.0]), np.array([0.0, 0.5]), np.array([0.25, 0.5])):
        g = np.ones((len(tv), 1))
        r = emu.adjoint(0.0, tv, np.ones((2, 1)), np.full((2, 1), 2.0), g, 1e-8, 1e-8, hist_cap=128)
        yo, go, lo, so, _ = orc.solve_adjoint(0.0, tv, np.ones((2, 1)), np.full((2, 1), 2.0), g)
        assert (r['status'] == 0).all() and (so == 0).all()
        np.testing.assert_allclose(r['y'], yo, rtol=1e-9)
        np.testing.assert_allclose(r['lamda'], lo, rtol=1e-9)


def test_ten_state_chain_uses_loop_lu(tmp_path):
    """A 10-state linear reaction chain: exercises the loop-based (run-time indexed) LU that the
    one-lane-per-instance build uses beyond 8 states (forward kernel, SB_NO_GROUP backward), and the
    grouped backward kernel with 8 lanes x 2 components of which 6 are padding (one state and one
    parameter: a single quadrature component for 8 lanes); both against the oracle."""
    from sunode_b200 import SympyProblem

    def rhs(t, y, p):
        x = y.x
        out = []
        for i in range(10):
            inflow = p.k * x[i - 1] if i > 0 else 0
            out.append(inflow - p.k * (1 + 0.1 * i) * x[i])
        return {'x': out}

    prob = SympyProblem({'k': ()}, {'x': 10}, rhs, [('k',)])
    tv = np.linspace(0.2, 3.0, 8)
    y0 = np.zeros((4, 10)); y0[:, 0] = 1.0
    k = np.array([[0.5], [1.0], [2.0], [4.0]])
    g = np.random.default_rng(0).standard_normal((8, 10))
    emu = Emulator(prob, str(tmp_path), group=True)
    assert emu.lib.emu_group_size() == 8
    yo, go, lo, so, _ = Oracle(prob, rtol=1e-8, atol=1e-8).solve_adjoint(0.0, tv, y0, k, g)
    for group in (False, True):
        r = emu.adjoint(0.0, tv, y0, k, g, 1e-8, 1e-8, hist_cap=512, group=group)
        assert (r['status'] == 0).all() and (so == 0).all()
        assert np.max(np.abs(r['y'] - yo) / (1e-8 * np.abs(yo) + 1e-8)) <= 1e-2
        np.testing.assert_allclose(r['grad'], go, rtol=1e-7)
        np.testing.assert_allclose(r['lamda'], lo, rtol=1e-7, atol=1e-12)


@pytest.mark.parametrize('name', ['lv_adj', 'robertson_adj'])
def test_flat_interval_schedule_is_the_same_computation(name, tmp_path):
    """sb_backward_flat (every lane walks its intervals on its own) performs, per instance, exactly
    the operations of sb_backward (lanes restart together): bit-identical results and counters."""
    w = examples.workloads()[name]
    prob = w.make_problem()
    B = 8
    y0, theta = w.draws(B)
    grads = np.random.default_rng(5).standard_normal((len(w.tvals), prob.n_states))
    emu = Emulator(prob, str(tmp_path))
    a = emu.adjoint(w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=w.history_capacity)
    for flat in (0, 1, -1):
        b = emu.adjoint(w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=w.history_capacity,
                        flat=flat)
        for key in ('grad', 'lamda', 'status', 'stats'):
            np.testing.assert_array_equal(a[key], b[key])


def test_lane_group_code_matches_one_lane_per_instance(tmp_path):
    """csrc/sb_group.cuh on the host: the lanes of a SEIR instance are threads, shuffles and
    __syncwarp are barrier rendezvous (tests/emu/cuda_shim_group.h); checked for 4 lanes x 2 state
    components (the default), 8 x 1 and 2 x 4.  The cross-lane LU, the
    butterfly norms and the shared-memory exchange must reproduce the one-lane-per-instance
    integrator: same step sequence but for rounding-level flips, gradients to 1e-9."""
    w = examples.workloads()['seir_adj']
    prob = w.make_problem()
    B = 3
    y0, theta = w.draws(B)
    grads = np.random.default_rng(12).standard_normal((B, len(w.tvals), prob.n_states))
    one = None
    for lanes in (None, 8, 2):
        defines = () if lanes is None else ('SB_GROUP_LANES=%d' % lanes,)
        emu = Emulator(prob, str(tmp_path), defines=defines, group=True)
        assert emu.lib.emu_group_size() == (lanes or 4)
        if one is None:
            one = emu.adjoint(w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=512)
        grp = emu.adjoint(w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=512, group=True)
        assert (one['status'] == 0).all() and (grp['status'] == 0).all()
        assert np.max(np.abs(grp['grad'] - one['grad']) / np.abs(one['grad']).max(axis=0)) <= 1e-9
        assert np.max(np.abs(grp['lamda'] - one['lamda']) / np.abs(one['lamda']).max(axis=0)) <= 1e-9
        steps_one, steps_grp = one['stats'][:, 0], grp['stats'][:, 0]
        assert np.max(np.abs(steps_one - steps_grp) / steps_one) <= 0.01
        np.testing.assert_array_equal(one['stats'][:, 7], grp['stats'][:, 7])      # stored points


def test_forward_lane_groups_match_one_lane_per_instance(tmp_path):
    """The grouped forward drivers (csrc/sb_group.cuh: forward_instance_group, plain and with
    forward sensitivities) on the host, lanes as threads: trajectories, sensitivities and the
    stored step history of the SEIR problem against the one-lane integrator -- the same step
    sequence but for rounding-level flips, values to 1e-3 tolerance units."""
    w = examples.workloads()['seir_adj']
    prob = w.make_problem()
    B = 3
    y0, theta = w.draws(B)
    for lanes in (None, 8):
        defines = () if lanes is None else ('SB_GROUP_LANES=%d' % lanes,)
        emu = Emulator(prob, str(tmp_path), defines=defines, group=True)
        one = emu.forward(w.t0, w.tvals, y0, theta, 1e-8, 1e-8, hist_cap=512)
        grp = emu.forward(w.t0, w.tvals, y0, theta, 1e-8, 1e-8, hist_cap=512, group=True)
        assert (one['status'] == 0).all() and (grp['status'] == 0).all()
        tol = 1e-8 * np.abs(one['y']) + 1e-8
        assert np.max(np.abs(grp['y'] - one['y']) / tol) <= 1e-3
        assert np.max(np.abs(one['stats'][:, 0] - grp['stats'][:, 0]) / one['stats'][:, 0]) <= 0.01
        np.testing.assert_array_equal(one['hist_n'], grp['hist_n'])
        for b in range(B):                               # the stored (t, order, y) points
            n = one['hist_n'][b]
            np.testing.assert_allclose(grp['hist'][b, :n], one['hist'][b, :n], rtol=1e-9, atol=1e-12)
    s0 = np.zeros((prob.n_params, prob.n_states))
    one = emu.forward_sens(w.t0, w.tvals, y0, theta, s0, 1e-6, 1e-6)
    grp = emu.forward_sens(w.t0, w.tvals, y0, theta, s0, 1e-6, 1e-6, group=True)
    assert (one['status'] == 0).all() and (grp['status'] == 0).all()
    assert np.max(np.abs(grp['y'] - one['y']) / (1e-6 * np.abs(one['y']) + 1e-6)) <= 1e-3
    scale = np.abs(one['sens']).max(axis=(0, 1))
    assert np.max(np.abs(grp['sens'] - one['sens']) / np.maximum(scale, 1e-300)) <= 1e-7
    assert np.max(np.abs(one['stats'][:, 0] - grp['stats'][:, 0]) / one['stats'][:, 0]) <= 0.02


def test_double_integrator_zero_error_estimates(tmp_path):
    """SURVEY G2 (from_sympy.ipynb cells 39-41) through the device code: the solution is a
    quadratic, the local error estimates are exactly zero from order 2 on, so the step-size ratio
    takes its zero / out-of-range branch (eta_root2's slow path) and the step grows by etamax."""
    from tests.test_oracle import _double_integrator, _double_integrator_closed_form
    prob = _double_integrator()
    tvals = np.arange(1, 10).astype(float)
    rng = np.random.default_rng(41)
    B = 4
    P, Y0 = rng.standard_normal((B, 3)), rng.standard_normal((B, 2))
    emu = Emulator(prob, str(tmp_path))
    fwd = emu.forward(0.0, tvals, Y0, P, 1e-10, 1e-10)
    grads = 2 * fwd['y']
    r = emu.adjoint(0.0, tvals, Y0, P, grads, 1e-10, 1e-10, hist_cap=512)
    assert (r['status'] == 0).all()
    for b in range(B):
        sol, loss, grad_p, grad_y0 = _double_integrator_closed_form(tvals, Y0[b], P[b])
        np.testing.assert_allclose(r['y'][b], sol, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(r['grad'][b], grad_p, rtol=1e-8, atol=1e-8 * abs(grad_p[1]))
        np.testing.assert_allclose(-r['lamda'][b], grad_y0, rtol=1e-8, atol=1e-8 * np.abs(grad_y0).max())


def test_argument_blocks_match_their_ctypes_mirrors(tmp_path):
    """csrc/sb_args.h is shared by the launcher and the device code; the emulation reaches the
    device code through ctypes mirrors of those structs, which must not drift."""
    import ctypes
    from tests.emu import emu as E
    em = Emulator(examples.lotka_volterra(), str(tmp_path))
    sizes = [ctypes.c_int() for _ in range(4)]
    em.lib.emu_arg_sizes(*[ctypes.byref(s) for s in sizes])
    assert sizes[0].value == ctypes.sizeof(E.ForwardArgs)
    assert sizes[1].value == ctypes.sizeof(E.TablesArgs)
    assert sizes[2].value == ctypes.sizeof(E.BackwardArgs)
