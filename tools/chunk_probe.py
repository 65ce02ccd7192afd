#!/usr/bin/env python
"""Bounded memory at scale (GPU box): Robertson fwd+adjoint at 262 144 draws -- history + tables of
one launch would be 288 GB -- through sb_solve_adjoint's chunking; compares a sample of the
results with the same draws solved in one small launch."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from sunode_b200 import examples  # noqa: E402
from sunode_b200.solver import AdjointSolver  # noqa: E402

w = examples.workloads()['robertson_adj']
prob = w.make_problem()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
rng = np.random.default_rng(1)
y0 = np.tile(np.asarray(w.y0, float), (B, 1))
theta = np.asarray(w.theta_med) * np.exp(w.sigma * rng.standard_normal((B, 3)))
g = w.grads(3)
solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=w.history_capacity)
for label in ('first call (allocates the stores)', 'second call'):
    t = time.perf_counter()
    y, grad, lam, st = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, g)
    dt = time.perf_counter() - t
    print('B = %d, %s: %d chunks, %.2f s, %.3e solves/s end to end from host arrays, failed %d'
          % (B, label, solver._engine.last_chunks(), dt, B / dt, int((st != 0).sum())))
idx = np.linspace(0, B - 1, 512).astype(int)
small = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=w.history_capacity)
y2, g2, l2, s2 = small.solve_adjoint_batch(w.t0, w.tvals, y0[idx], theta[idx], g)
ok = (st[idx] == 0) & (s2 == 0)
print('sample of 512 against a single small launch: identical y %s, identical grad %s, identical lamda %s'
      % (np.array_equal(y[idx][ok], y2[ok]), np.array_equal(grad[idx][ok], g2[ok]), np.array_equal(lam[idx][ok], l2[ok])))
