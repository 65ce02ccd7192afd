#!/usr/bin/env python
"""Host-side cost of one asynchronous (device-memory) sb_solve_adjoint call: wall time of the call
itself against the device time of the step (GPU box)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402,F401
import torch
from sunode_b200 import examples
from sunode_b200.solver import AdjointSolver

w = examples.workloads()['lv_adj']
prob = w.make_problem()
B = w.batch
dev = torch.device('cuda', 0)
y0, theta = (torch.from_numpy(a).to(dev) for a in w.draws(B))
grads = torch.from_numpy(w.grads(2)).to(dev)
s = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512)
out = dict(y_out=torch.empty((B, 50, 2), dtype=torch.float64, device=dev),
           grad_out=torch.empty((B, 2), dtype=torch.float64, device=dev),
           lamda_out=torch.empty((B, 2), dtype=torch.float64, device=dev),
           status=torch.empty((B,), dtype=torch.int32, device=dev))
for _ in range(3):
    s.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads, **out)
torch.cuda.synchronize()
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    e0.record()
    s.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads, **out)
    e1.record()
    host = time.perf_counter() - t
    torch.cuda.synchronize()
    print('host call %.3f ms   device span %.3f ms   kernels %s' % (1e3 * host, e0.elapsed_time(e1), s._engine.last_kernel_ms()))
