"""GPU tier, needs >= 2 devices (skipped otherwise): the sharded solve over NCCL equals the
single-GPU solve of the global batch, bit for bit (instances are independent)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmpdir, B):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from sunode_b200 import examples
    from sunode_b200.sharding import solve_adjoint_sharded
    from sunode_b200.solver import AdjointSolver
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank,
                            world_size=world, device_id=torch.device('cuda', rank))
    try:
        w = examples.workloads()['lv_adj']
        prob = w.make_problem()
        y0, theta = w.draws(B)
        grads = np.random.default_rng(5).standard_normal((B, len(w.tvals), prob.n_states))
        dev = torch.device('cuda', rank)
        solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512, device=rank)
        y, g, lam, st = solve_adjoint_sharded(
            solver, w.t0, w.tvals, torch.from_numpy(y0).to(dev), torch.from_numpy(theta).to(dev),
            torch.from_numpy(grads).to(dev))
        torch.cuda.synchronize()
        np.savez(os.path.join(tmpdir, 'out%d.npz' % rank), y=y.cpu().numpy(), g=g.cpu().numpy(),
                 lam=lam.cpu().numpy(), st=st.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_two_gpu_nccl_equals_single_gpu(tmp_path):
    torch = pytest.importorskip('torch')
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from sunode_b200 import examples
    from sunode_b200.solver import AdjointSolver
    B = 1000                                      # not a multiple of 2 * 32: ragged shards
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), B), nprocs=2, join=True)
    w = examples.workloads()['lv_adj']
    prob = w.make_problem()
    y0, theta = w.draws(B)
    grads = np.random.default_rng(5).standard_normal((B, len(w.tvals), prob.n_states))
    solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512, device=0)
    y, g, lam, st = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads)
    for rank in range(2):
        out = np.load(tmp_path / ('out%d.npz' % rank))
        np.testing.assert_array_equal(out['y'], y)
        np.testing.assert_array_equal(out['g'], g)
        np.testing.assert_array_equal(out['lam'], lam)
        np.testing.assert_array_equal(out['st'], st)
