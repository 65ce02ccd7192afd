"""Failure statistics of a workload over several batches of draws (diagnostics)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sunode_b200 import examples
from sunode_b200.solver import AdjointSolver

name = sys.argv[1] if len(sys.argv) > 1 else 'robertson_adj'
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = examples.workloads()[name]
prob = w.make_problem()
g = w.grads(prob.n_states)
solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=w.history_capacity)
tot = 0
codes = {}
for b in range(nb):
    y0, theta = w.draws(w.batch, offset=b * w.batch)
    y, gr, lam, st = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, g)
    bad = np.nonzero(st)[0]
    tot += len(bad)
    for c in st[bad]:
        codes[int(c)] = codes.get(int(c), 0) + 1
    print('batch', b, 'failed', bad.tolist(), st[bad].tolist(), flush=True)
print('defines', os.environ.get('SUNODE_B200_DEFINES', ''), 'total failed', tot, 'of', nb * w.batch, codes)
