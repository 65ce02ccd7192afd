"""Find instances of a workload that fail on the GPU and re-run them alone (diagnostics)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sunode_b200 import examples  # noqa: E402
from sunode_b200.solver import AdjointSolver  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'robertson_adj'
    w = examples.workloads()[name]
    prob = w.make_problem()
    y0, theta = w.draws()
    g = w.grads(prob.n_states)
    solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=w.history_capacity)
    B = len(y0)
    sf = np.zeros((B, 8), np.int32)
    sb = np.zeros((B, 8), np.int32)
    y, gr, lam, st = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, g, stats_fwd=sf, stats_bwd=sb)
    bad = np.nonzero(st)[0]
    print('defines', os.environ.get('SUNODE_B200_DEFINES', ''), 'failed', bad, st[bad])
    for i in bad:
        print('inst', i, 'theta', repr(theta[i]), 'fwd stats', sf[i], 'bwd stats', sb[i])
    idx = [int(a) for a in sys.argv[2:]] or list(bad)
    for i in idx:
        sf1 = np.zeros((1, 8), np.int32)
        sb1 = np.zeros((1, 8), np.int32)
        y1, g1, l1, s1 = solver.solve_adjoint_batch(w.t0, w.tvals, y0[i:i + 1], theta[i:i + 1], g,
                                                    stats_fwd=sf1, stats_bwd=sb1)
        print('alone inst', i, 'status', s1, 'fwd', sf1[0], 'bwd', sb1[0], 'grad', g1[0])


if __name__ == '__main__':
    main()
