"""Small adjoint solves of every benchmark problem (run under compute-sanitizer --tool memcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sunode_b200 import examples
from sunode_b200.solver import AdjointSolver, Solver
ONLY = os.environ.get('PROBE_ONLY')
for name, B in (('lv_adj', 70), ('robertson_adj', 37), ('seir_adj', 21)):
    if ONLY and name != ONLY:
        continue
    w = examples.workloads()[name]
    prob = w.make_problem()
    y0, theta = w.draws(B)
    g = w.grads(prob.n_states)
    for seg in ('1', '5'):
        os.environ['SUNODE_B200_SEGMENTS'] = seg
        s = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=w.history_capacity)
        out = s.solve_adjoint_batch(w.t0, w.tvals, y0, theta, g)
        print(name, 'segments', seg, 'failed', int((out[3] != 0).sum()), 'grad[0]', out[1][0])
    f = Solver(prob, abstol=1e-8, reltol=1e-8, sens_mode='simultaneous')
    ys, ss, st = f.solve_sens_batch(w.t0, w.tvals, y0, theta, np.zeros((prob.n_params, prob.n_states)))
    print(name, 'forward sens: failed', int((st != 0).sum()), 'sens[0, -1, 0]', ss[0, -1, 0])
    if name == 'lv_adj':
        os.environ['SUNODE_B200_SEGMENTS'] = '1'
        fs = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512, backward='fundamental')
        print(name, 'restart-free pass: failed', int((fs.solve_adjoint_batch(w.t0, w.tvals, y0, theta, g)[3] != 0).sum()))
