#!/usr/bin/env python
"""Where on the device timeline does the peer-to-peer trajectory gather run?  (torchrun, N >= 2)
Prints, per step, the time from the step's start to: forward done, gather copies done, backward done."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from sunode_b200 import examples, sharding  # noqa: E402
from sunode_b200.solver import AdjointSolver  # noqa: E402

rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
w = examples.workloads()['lv_adj']
prob = w.make_problem()
B = w.batch
y0, th = (torch.from_numpy(a).to(dev) for a in w.draws(B, offset=rank * B))
grads = torch.from_numpy(w.grads(2)).to(dev)
solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512, device=lr)
sym, handle = sharding.symmetric_rows((B, 50, 2), dev)
y_all = torch.empty((world * B, 50, 2), dtype=torch.float64, device=dev)
main = torch.cuda.current_stream()
for it in range(6):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    y, _ = solver.solve_forward_batch(w.t0, w.tvals, y0, th, y_out=sym)
    ev[1].record()
    g, lam, st = solver.solve_backward_batch(w.tvals[-1], w.t0, w.tvals, grads)
    ev[3].record()
    pg = sharding._PeerGather(y, handle, world, ev[1], out=y_all)
    with torch.cuda.stream(pg.side):
        ev[2].record()
    pg.result()
    torch.cuda.synchronize()
    dist.barrier()
    if it >= 3:
        print('rank %d step %d: forward done %.3f ms, gather copies done %.3f ms, backward done %.3f ms'
              % (rank, it, ev[0].elapsed_time(ev[1]), ev[0].elapsed_time(ev[2]), ev[0].elapsed_time(ev[3])), flush=True)
dist.destroy_process_group()
