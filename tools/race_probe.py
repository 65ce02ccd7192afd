"""Determinism probe of the grouped-lane backward kernel (A/B tool)."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sunode_b200 import examples
from sunode_b200.solver import AdjointSolver
w = examples.workloads()['seir_adj']
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
y0, theta = w.draws(B)
grads = np.random.default_rng(8).standard_normal((B, len(w.tvals), 8))
for trial in range(3):
    solver = AdjointSolver(w.make_problem(), abstol=1e-8, reltol=1e-8, history_capacity=512)
    ref = None
    for it in range(n):
        seg = (None, '1', '7', '51', '1', None)[it % 6]
        if seg is None: os.environ.pop('SUNODE_B200_SEGMENTS', None)
        else: os.environ['SUNODE_B200_SEGMENTS'] = seg
        sb = np.zeros((B, 8), np.int32)
        out = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads, stats_bwd=sb)
        if ref is None: ref = (out, sb.copy())
        else:
            bad = (sb != ref[1]).any(axis=1)
            print(trial, it, 'seg', seg, 'grad equal', np.array_equal(out[1], ref[0][1]), 'rows', np.nonzero(bad)[0], (sb - ref[1])[bad])
