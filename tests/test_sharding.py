"""Host logic of the multi-GPU path on CPU: world_size-2 gloo processes, each solving its shard
(the device integrator is stood in for by its host emulation, tests/emu) and all-gathering the
outputs; the result must equal the single-process solve of the global batch."""
import os
import sys

import numpy as np
import pytest

from sunode_b200.sharding import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


class _EmuSolver:
    """Stand-in with the AdjointSolver batch signature, computing with the emulated device code."""

    def __init__(self, problem, workdir):
        from tests.emu.emu import Emulator
        self.emu = Emulator(problem, workdir)

    def solve_adjoint_batch(self, t0, tvals, y0, params, grads):
        r = self.emu.adjoint(t0, tvals, np.asarray(y0), np.asarray(params), np.asarray(grads),
                             1e-8, 1e-8, hist_cap=512)
        return r['y'], r['grad'], r['lamda'], r['status']

    # the split calls (what the overlapped path of sharding.solve_adjoint_gathered uses)
    def solve_forward_batch(self, t0, tvals, y0, params, y_out=None):
        self._fwd = (t0, np.asarray(y0), np.asarray(params))
        r = self.emu.forward(t0, tvals, np.asarray(y0), np.asarray(params), 1e-8, 1e-8)
        return r['y'], r['status']

    def solve_backward_batch(self, t_last, t_first, tvals, grads, grad_out=None, lamda_out=None,
                             status=None):
        t0, y0, params = self._fwd
        assert t_first == t0 and t_last == tvals[-1]
        r = self.emu.adjoint(t0, tvals, y0, params, np.asarray(grads), 1e-8, 1e-8, hist_cap=512)
        return r['grad'], r['lamda'], r['status']


def _worker(rank, world, port, tmpdir, B, overlap):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from sunode_b200 import examples
    from sunode_b200.sharding import solve_adjoint_sharded
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank,
                            world_size=world)
    try:
        w = examples.workloads()['lv_adj']
        prob = w.make_problem()
        y0, theta = w.draws(B)
        rng = np.random.default_rng(5)
        grads = rng.standard_normal((B, len(w.tvals), prob.n_states))
        solver = _EmuSolver(prob, os.path.join(tmpdir, 'emu%d' % rank))
        y, g, lam, st = solve_adjoint_sharded(solver, w.t0, w.tvals, y0, theta, grads,
                                              overlap=overlap)
        np.savez(os.path.join(tmpdir, 'out%d.npz' % rank), y=y, g=g, lam=lam, st=st)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('B,overlap', [(37, False), (64, True), (37, True)])
def test_two_rank_gloo_equals_single_process(tmp_path, B, overlap):
    """``overlap``: forward, trajectories' all-gather in flight (async), backward, small gather --
    the order the GPU path uses to hide the collective under the backward kernels."""
    import socket
    import torch.multiprocessing as mp
    from sunode_b200 import examples
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), B, overlap), nprocs=2, join=True)

    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    y0, theta = w.draws(B)
    grads = np.random.default_rng(5).standard_normal((B, len(w.tvals), prob.n_states))
    ref = _EmuSolver(prob, str(tmp_path / 'emu_ref')).solve_adjoint_batch(
        w.t0, w.tvals, y0, theta, grads)
    for rank in range(2):
        out = np.load(tmp_path / ('out%d.npz' % rank))
        np.testing.assert_array_equal(out['y'], ref[0])
        np.testing.assert_array_equal(out['g'], ref[1])
        np.testing.assert_array_equal(out['lam'], ref[2])
        np.testing.assert_array_equal(out['st'], ref[3])
