"""Experiment: does grouping instances of similar cost into the same warp pay?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sunode_b200 import examples
from sunode_b200.solver import AdjointSolver

w = examples.workloads()['lv_adj']
prob = w.make_problem()
y0, theta = w.draws()
g = w.grads(2)
solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512)
B = len(y0)
sf = np.zeros((B, 8), np.int32); sb = np.zeros((B, 8), np.int32)
solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, g, stats_fwd=sf, stats_bwd=sb)
dev = torch.device('cuda:0')

def timeit(order, label):
    y0d = torch.from_numpy(np.ascontiguousarray(y0[order])).to(dev)
    thd = torch.from_numpy(np.ascontiguousarray(theta[order])).to(dev)
    gd = torch.from_numpy(g).to(dev)
    outs = None
    for _ in range(3):
        outs = solver.solve_adjoint_batch(w.t0, w.tvals, y0d, thd, gd)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        solver.solve_adjoint_batch(w.t0, w.tvals, y0d, thd, gd)
        torch.cuda.synchronize()
        ts.append(solver._engine.last_kernel_ms())
    ts = np.array(ts).mean(axis=0)
    print('%-28s fwd %.3f tab %.3f bwd %.3f ms' % (label, *ts))

timeit(np.arange(B), 'unsorted')
timeit(np.argsort(sf[:, 0], kind='stable'), 'sorted by fwd steps')
timeit(np.argsort(sb[:, 0], kind='stable'), 'sorted by bwd steps (ideal)')
timeit(np.argsort(theta[:, 0] + theta[:, 3], kind='stable'), 'sorted by alpha+delta')
timeit(np.argsort(-sb[:, 0], kind='stable'), 'sorted by bwd steps desc')
