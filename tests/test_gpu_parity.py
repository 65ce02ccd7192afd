"""GPU parity tests (``-m gpu``): the CUDA path, called through the C ABI exactly as a user
would (``Solver`` / ``AdjointSolver``), against the CPU oracle on the same seeded inputs.

Tolerances.  Both sides implement the same algorithm (CVODES-style BDF) in double precision but
with different operation order / FMA contraction, so they agree to rounding level only as long as
they take the same step sequence; for non-stiff problems they do (difference << 1 tolerance
unit), for the stiff Robertson problem the sequences decorrelate and each side is only within the
solver's *global* error of the truth.  The tests therefore check
  (i)  GPU vs oracle:   |y_gpu - y_cpu| <= ENV * (rtol*|y| + atol)   with ENV stated per problem,
  (ii) gradients:       rel. error <= 1e-5 at rtol = atol = 1e-8 (backward tolerances 1e-10, as
                        the reference hard-codes, solver.py:599,614), measured against the
                        column scale of the batch.
"""
import numpy as np
import pytest

from sunode_b200 import examples
from sunode_b200.solver import AdjointSolver, Solver, SolverError

pytestmark = pytest.mark.gpu


def _oracle(problem, **kw):
    from oracle.oracle import Oracle
    return Oracle(problem, **kw)


CASES = {
    # name: (batch, trajectory envelope in tolerance units, gradient rtol)
    'lv_adj': (512, 1.0, 1e-7),
    'seir_adj': (256, 1.0, 1e-7),
    'robertson_adj': (128, 1000.0, 1e-5),
}


@pytest.mark.parametrize('name', list(CASES))
def test_forward_matches_oracle(name):
    B, env, _ = CASES[name]
    w = examples.workloads()[name]
    prob = w.make_problem()
    y0, theta = w.draws(B)
    solver = Solver(prob, abstol=1e-8, reltol=1e-8)
    stats = np.zeros((B, 8), dtype=np.int32)
    y, status = solver.solve_batch(w.t0, w.tvals, y0, theta, stats=stats)
    yo, so, sto = _oracle(prob, rtol=1e-8, atol=1e-8).solve_forward(w.t0, w.tvals, y0, theta)
    assert (status == 0).all() and (so == 0).all()
    tol = 1e-8 * np.abs(yo) + 1e-8
    assert np.max(np.abs(y - yo) / tol) <= env
    if env <= 1.0:
        # same step sequence => same counters (a rounding-level difference may flip a controller
        # decision for the odd instance; the trajectory envelope above still holds for it)
        assert (stats[:, 0] == sto[:, 0]).mean() >= 0.95
        assert (stats[:, 1] == sto[:, 1]).mean() >= 0.95


@pytest.mark.parametrize('name', list(CASES))
def test_adjoint_matches_oracle(name):
    B, env, grtol = CASES[name]
    w = examples.workloads()[name]
    prob = w.make_problem()
    y0, theta = w.draws(B)
    rng = np.random.default_rng(7)
    grads = rng.standard_normal((B, len(w.tvals), prob.n_states))
    solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=w.history_capacity)
    y, g, lam, status = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads)
    yo, go, lo, so, _ = _oracle(prob, rtol=1e-8, atol=1e-8).solve_adjoint(
        w.t0, w.tvals, y0, theta, grads)
    assert (status == 0).all() and (so == 0).all()
    tol = 1e-8 * np.abs(yo) + 1e-8
    assert np.max(np.abs(y - yo) / tol) <= env
    assert np.max(np.abs(g - go) / np.abs(go).max(axis=0)) <= grtol
    assert np.max(np.abs(lam - lo) / np.abs(lo).max(axis=0)) <= max(grtol, 1e-4 if env > 1 else grtol)


def test_separate_forward_backward_equals_fused():
    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    B = 96
    y0, theta = w.draws(B)
    grads = np.ones((len(w.tvals), prob.n_states))      # shared cotangent, test_solve.py:99
    solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512)
    y1, st1 = solver.solve_forward_batch(w.t0, w.tvals, y0, theta)
    g1, l1, sb1 = solver.solve_backward_batch(w.tvals[-1], w.t0, w.tvals, grads)
    y2, g2, l2, st2 = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads)
    assert (st1 == 0).all() and (sb1 == 0).all() and (st2 == 0).all()
    np.testing.assert_array_equal(y1, y2)
    np.testing.assert_array_equal(g1, g2)
    np.testing.assert_array_equal(l1, l2)


def test_reference_shaped_batch1_api():
    """The reference's smoke test (sunode/test_solve.py:81-154) with numbers attached:
    x' = x + b has the closed form x(t) = (1 + b) e^t - b."""
    from sunode_b200 import SympyProblem

    def rhs(t, y, p):
        return {'x': y.x + p.a.b}

    prob = SympyProblem({'a': {'b': ()}}, {'x': ()}, rhs, [('a', 'b')])
    b = 0.2
    time = np.linspace(0, 1)
    y0 = np.ones((1,), dtype=prob.state_dtype)[0]
    y0['x'] = 1.0

    solver = Solver(prob)
    solver.set_params_dict({'a': {'b': b}})
    out = solver.make_output_buffers(time)
    solver.solve(0, time, np.ones(1), out)
    np.testing.assert_allclose(out[:, 0], (1 + b) * np.exp(time) - b, rtol=1e-7)
    out2 = solver.make_output_buffers(time)
    solver.solve(0, time, y0, out2)           # structured scalar y0 (solver.py:489-490)
    np.testing.assert_array_equal(out, out2)
    with pytest.raises(ValueError):
        solver.solve(0, time, np.ones(2), out)

    adj = AdjointSolver(prob)
    adj.set_params_dict({'a': {'b': b}})
    y_out, grad_out, lamda_out = adj.make_output_buffers(time)
    adj.solve_forward(0, time, np.ones(1), y_out)
    grads = np.ones_like(y_out)
    adj.solve_backward(time[-1], 0, time, grads, grad_out, lamda_out)
    np.testing.assert_allclose(y_out[:, 0], (1 + b) * np.exp(time) - b, rtol=1e-7)
    np.testing.assert_allclose(grad_out[0], np.sum(np.exp(time) - 1), rtol=1e-6)
    np.testing.assert_allclose(-lamda_out[0], np.sum(np.exp(time)), rtol=1e-6)


def test_failed_instances_are_nan_with_cvodes_flag():
    """A draw that blows up (finite-time singularity) must not poison its neighbours: NaN rows and
    a CVODES flag for that instance only (as_pytensor.py:287-290 semantics, per instance)."""
    from sunode_b200 import SympyProblem

    def rhs(t, y, p):
        return {'x': p.k * y.x ** 2}

    prob = SympyProblem({'k': ()}, {'x': ()}, rhs, [('k',)])
    solver = Solver(prob, abstol=1e-8, reltol=1e-8)
    tvals = np.linspace(0.1, 2, 20)
    y0 = np.ones((3, 1))
    k = np.array([[0.1], [1.0], [0.2]])                  # x blows up at t = 1/k -> only k = 1 fails
    y, status = solver.solve_batch(0.0, tvals, y0, k)
    assert status[0] == 0 and status[2] == 0 and status[1] < 0
    assert np.isnan(y[1]).all()
    np.testing.assert_allclose(y[0, :, 0], 1 / (1 - 0.1 * tvals), rtol=1e-6)
    with pytest.raises(SolverError):
        solver.set_params(np.array((1.0,), dtype=prob.params_dtype)[()])
        solver.solve(0.0, tvals, np.ones(1), solver.make_output_buffers(tvals))


def test_device_tensors_run_in_place():
    torch = pytest.importorskip('torch')
    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    B = 256
    y0, theta = w.draws(B)
    grads = np.ones((len(w.tvals), prob.n_states))
    solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512)
    y_h, g_h, l_h, s_h = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads)
    dev = torch.device('cuda:0')
    y_d, g_d, l_d, s_d = solver.solve_adjoint_batch(
        w.t0, w.tvals, torch.from_numpy(y0).to(dev), torch.from_numpy(theta).to(dev),
        torch.from_numpy(grads).to(dev))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(y_d.cpu().numpy(), y_h)
    np.testing.assert_array_equal(g_d.cpu().numpy(), g_h)
    np.testing.assert_array_equal(l_d.cpu().numpy(), l_h)
    assert (s_d.cpu().numpy() == 0).all()


def test_full_size_properties_lv():
    """BASELINE config 3 at full size (B = 65 536): every instance succeeds, outputs are finite,
    identical draws give identical results (no cross-instance interference), and the gradient
    is linear in the cotangent (adjoint linearity)."""
    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    y0, theta = w.draws()
    theta[1::2] = theta[0::2]                            # pairs of identical draws
    g1 = np.ones((len(w.tvals), prob.n_states))
    solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512)
    y, g, lam, status = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, g1)
    assert (status == 0).all()
    assert np.isfinite(y).all() and np.isfinite(g).all() and np.isfinite(lam).all()
    np.testing.assert_array_equal(y[0::2], y[1::2])
    np.testing.assert_array_equal(g[0::2], g[1::2])
    _, g3, lam3, _ = solver.solve_adjoint_batch(w.t0, w.tvals, y0[:4096], theta[:4096], 3.0 * g1)
    np.testing.assert_allclose(g3, 3.0 * g[:4096], rtol=2e-6, atol=1e-9 * np.abs(g).max())
    np.testing.assert_allclose(lam3, 3.0 * lam[:4096], rtol=2e-6, atol=1e-9 * np.abs(lam).max())


def test_vector_atol_and_pickled_solver():
    """CVodeSVtolerances (solver.py:404-407) and Solver pickling (solver.py:319-324: the
    configuration travels, the engine handle is re-created)."""
    import pickle
    w = examples.workloads()['robertson_adj']
    prob = w.make_problem()
    B = 64
    y0, theta = w.draws(B)
    atol = np.array([1e-8, 1e-12, 1e-8])                 # the classic Robertson tolerances
    solver = Solver(prob, abstol=atol, reltol=1e-8)
    y, status = solver.solve_batch(w.t0, w.tvals, y0, theta)
    yo, so, _ = _oracle(prob, rtol=1e-8, atol=atol).solve_forward(w.t0, w.tvals, y0, theta)
    assert (status == 0).all() and (so == 0).all()
    tol = 1e-8 * np.abs(yo) + atol
    assert np.max(np.abs(y - yo) / tol) <= 1000.0
    clone = pickle.loads(pickle.dumps(solver))
    y2, status2 = clone.solve_batch(w.t0, w.tvals, y0, theta)
    np.testing.assert_array_equal(y, y2)


def test_edge_cases_and_backward_traces():
    """Batch sizes that do not fill a warp, all output times equal to t0, and the optional
    lamda_all_out / quad_all_out traces of solve_backward (solver.py:723-724, 778-781)."""
    from sunode_b200 import SympyProblem
    prob = SympyProblem({'k': ()}, {'x': ()}, lambda t, y, p: {'x': -p.k * y.x}, [('k',)])
    solver = AdjointSolver(prob, abstol=1e-10, reltol=1e-10)
    for B in (1, 33):
        k = np.full((B, 1), 2.0)
        y, g, lam, st = solver.solve_adjoint_batch(0.0, np.array([0.0]), np.ones((B, 1)), k,
                                                  np.ones((1, 1)))
        assert (st == 0).all()
        np.testing.assert_array_equal(y[:, 0, 0], 1.0)
        np.testing.assert_array_equal(lam[:, 0], -1.0)
        np.testing.assert_array_equal(g[:, 0], 0.0)
    # traces: x = exp(-k t); after the jump at t_i (walking backward) lamda = -sum_{j>=i} e^{-k(t_j-t_i)}
    tv = np.array([0.5, 1.0, 1.5])
    solver.set_params(np.array((2.0,), dtype=prob.params_dtype)[()])
    y_out, grad_out, lamda_out = solver.make_output_buffers(tv)
    solver.solve_forward(0.0, tv, np.ones(1), y_out)
    lam_all, quad_all = np.zeros((3, 1)), np.zeros((3, 1))
    solver.solve_backward(tv[-1], 0.0, tv, np.ones((3, 1)), grad_out, lamda_out,
                          lamda_all_out=lam_all, quad_all_out=quad_all)
    expect = {2: -1.0, 1: -(1 + np.exp(-1.0)), 0: -(1 + np.exp(-1.0) + np.exp(-2.0))}
    # row convention of the reference: jump number i (0 = last time) lands in row (-i) % n_t
    np.testing.assert_allclose(lam_all[0, 0], expect[2], rtol=1e-8)
    np.testing.assert_allclose(lam_all[2, 0], expect[1], rtol=1e-8)
    np.testing.assert_allclose(lam_all[1, 0], expect[0], rtol=1e-8)
    np.testing.assert_allclose(-lamda_out[0], (np.exp(-1.0) + np.exp(-2.0) + np.exp(-3.0)), rtol=1e-8)
    np.testing.assert_allclose(grad_out[0], np.sum(-tv * np.exp(-2.0 * tv)), rtol=1e-7)
    assert quad_all[0, 0] == 0.0 and abs(quad_all[1, 0] - grad_out[0]) < abs(grad_out[0])


def test_segmented_work_queue_is_transparent(monkeypatch):
    """The backward pass cut into (group, segment) work units must give bit-identical results
    whatever the segmentation, including one unit per interval and batches smaller than a warp."""
    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512)
    rng = np.random.default_rng(3)
    for B in (5, 2048 + 17):
        y0, theta = w.draws(B)
        grads = rng.standard_normal((B, len(w.tvals), prob.n_states))
        ref = None
        for seg in ('1', '3', '10', '51', '1000'):
            monkeypatch.setenv('SUNODE_B200_SEGMENTS', seg)
            sb = np.zeros((B, 8), np.int32)
            out = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads, stats_bwd=sb)
            assert (out[3] == 0).all()
            if ref is None:
                ref = out + (sb,)
            else:
                for a, b in zip(ref, out + (sb,)):
                    np.testing.assert_array_equal(a, b)


def test_interval_schedule_is_transparent(monkeypatch):
    """sb_backward (lanes of a warp restart together) and sb_backward_flat (every lane walks its
    intervals on its own) perform the same per-instance computation: bit-identical outputs and
    counters, whichever the device-side rule or SUNODE_B200_FLAT picks (stiff and non-stiff)."""
    for name, B in (('lv_adj', 100), ('robertson_adj', 70)):
        w = examples.workloads()[name]
        prob = w.make_problem()
        solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=w.history_capacity)
        y0, theta = w.draws(B)
        grads = np.random.default_rng(4).standard_normal((B, len(w.tvals), prob.n_states))
        ref = None
        for flat in (None, '0', '1'):
            if flat is None:
                monkeypatch.delenv('SUNODE_B200_FLAT', raising=False)
            else:
                monkeypatch.setenv('SUNODE_B200_FLAT', flat)
            sb = np.zeros((B, 8), np.int32)
            out = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads, stats_bwd=sb) + (sb,)
            assert (out[3] == 0).all()
            if ref is None:
                ref = out
            else:
                for a, b in zip(ref, out):
                    np.testing.assert_array_equal(a, b)


def test_lane_groups_match_one_lane_per_instance(monkeypatch):
    """SEIR (8 states) runs with 4 lanes per instance (sb_group.cuh: two state components and
    matrix rows per lane, LU across the lanes with shuffles, norms as butterfly sums).  Same algorithm as the
    one-lane-per-instance build (-DSB_NO_GROUP), different summation order in the norms: the step
    sequences agree but for rounding-level decision flips and the results to 1e-9; segmentation
    is bit-transparent in group mode too; batch sizes that leave groups / warps partly empty."""
    w = examples.workloads()['seir_adj']
    prob = w.make_problem()
    rng = np.random.default_rng(8)
    for B in (3, 4 * 37 + 1):
        y0, theta = w.draws(B)
        grads = rng.standard_normal((B, len(w.tvals), prob.n_states))
        outs = {}
        for defines in ('', 'SB_NO_GROUP', 'SB_GROUP_CHECK'):
            monkeypatch.setenv('SUNODE_B200_DEFINES', defines)
            solver = AdjointSolver(w.make_problem(), abstol=1e-8, reltol=1e-8, history_capacity=512)
            sb = np.zeros((B, 8), np.int32)
            out = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads, stats_bwd=sb) + (sb,)
            assert (out[3] == 0).all()
            outs[defines] = out
            if defines in ('', 'SB_GROUP_CHECK'):
                for seg in ('1', '7', '51'):
                    monkeypatch.setenv('SUNODE_B200_SEGMENTS', seg)
                    sb2 = np.zeros((B, 8), np.int32)
                    out2 = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads, stats_bwd=sb2) + (sb2,)
                    for a, b in zip(out, out2):
                        np.testing.assert_array_equal(a, b)
                monkeypatch.delenv('SUNODE_B200_SEGMENTS')
        # the build that verifies, on the device, that the lanes of a group ran together wherever
        # they update their shared per-instance state (status -1006 otherwise)
        for a, b in zip(outs[''], outs['SB_GROUP_CHECK']):
            np.testing.assert_array_equal(a, b)
        g, t = outs[''], outs['SB_NO_GROUP']
        np.testing.assert_array_equal(g[0], t[0])                      # forward: same kernel
        scale = np.abs(t[1]).max(axis=0)
        assert np.max(np.abs(g[1] - t[1]) / scale) <= 1e-9
        assert np.max(np.abs(g[2] - t[2]) / np.abs(t[2]).max(axis=0)) <= 1e-9
        if B > 100:
            assert (g[4][:, 0] == t[4][:, 0]).mean() >= 0.9           # backward step counts


def test_ten_state_chain_grouped_lanes_with_padding():
    """10 states, 1 parameter: the grouped backward kernel with 8 lanes x 2 components (6 padding
    rows, one quadrature component for 8 lanes) and the one-lane forward kernel with the loop-based
    LU, against the oracle; a batch that leaves groups and warps partly empty."""
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
    B = 11
    y0 = np.zeros((B, 10)); y0[:, 0] = 1.0
    k = np.linspace(0.4, 4.0, B)[:, None]
    g = np.random.default_rng(0).standard_normal((8, 10))
    solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512)
    y, grad, lam, st = solver.solve_adjoint_batch(0.0, tv, y0, k, g)
    yo, go, lo, so, _ = _oracle(prob, rtol=1e-8, atol=1e-8).solve_adjoint(0.0, tv, y0, k, g)
    assert (st == 0).all() and (so == 0).all()
    assert np.max(np.abs(y - yo) / (1e-8 * np.abs(yo) + 1e-8)) <= 1e-2
    np.testing.assert_allclose(grad, go, rtol=1e-7)
    np.testing.assert_allclose(lam, lo, rtol=1e-7, atol=1e-12)


def test_chunked_adjoint_equals_single_launch():
    """Bounded memory (`AdjointSolver.set_workspace_limit`, the counterpart of the reference's
    `checkpoint_n`, solver.py:533,588): a workspace limit below the batch's history + tables cuts
    `sb_solve_adjoint` into chunks; host arrays and device tensors, results bit for bit those of
    the single launch."""
    torch = pytest.importorskip('torch')
    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    B = 5000
    y0, theta = w.draws(B)
    grads = np.random.default_rng(11).standard_normal((B, len(w.tvals), prob.n_states))
    one = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512)
    ref = one.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads)
    assert one._engine.last_chunks() == 1 and (ref[3] == 0).all()
    per_instance = 512 * ((2 + 2) + (10 + 6 * 2)) * 8
    cut = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512)
    cut.set_workspace_limit(2048 * per_instance + 1)
    out = cut.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads)
    assert cut._engine.last_chunks() == 3                      # 2048 + 2048 + 904
    for a, b in zip(ref, out):
        np.testing.assert_array_equal(a, b)
    dev = torch.device('cuda', 0)
    td = [torch.from_numpy(a).to(dev) for a in (y0, theta, grads)]
    yd, gd, ld, sd = cut.solve_adjoint_batch(w.t0, w.tvals, *td)
    torch.cuda.synchronize()
    assert cut._engine.last_chunks() == 3
    for a, b in zip(ref, (yd, gd, ld, sd)):
        np.testing.assert_array_equal(a, b.cpu().numpy())
