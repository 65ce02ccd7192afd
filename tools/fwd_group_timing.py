#!/usr/bin/env python
"""SEIR forward-only and forward-sensitivity throughput, lane groups against one lane per instance
(GPU box).  usage: tools/fwd_group_timing.py [batch]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from sunode_b200 import examples  # noqa: E402
from sunode_b200.solver import Solver  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
w = examples.workloads()['seir_adj']
prob = w.make_problem()
dev = torch.device('cuda', 0)
y0, th = (torch.from_numpy(a).to(dev) for a in w.draws(B))
for defs in ('', 'SB_NO_FWD_GROUP', 'SB_GROUP_MIN_BLOCKS=8'):
    if defs:
        os.environ['SUNODE_B200_DEFINES'] = defs
    else:
        os.environ.pop('SUNODE_B200_DEFINES', None)
    for sens in (False, True):
        solver = Solver(prob, abstol=1e-8, reltol=1e-8, sens_mode='simultaneous' if sens else None)
        y = torch.empty((B, 50, 8), dtype=torch.float64, device=dev)
        st = torch.empty((B,), dtype=torch.int32, device=dev)
        if sens:
            s0 = torch.zeros((6, 8), dtype=torch.float64, device=dev)
            so = torch.empty((B, 50, 6, 8), dtype=torch.float64, device=dev)
            call = lambda: solver.solve_sens_batch(w.t0, w.tvals, y0, th, s0, y_out=y, sens_out=so, status=st)  # noqa: E731
        else:
            call = lambda: solver.solve_batch(w.t0, w.tvals, y0, th, y_out=y, status=st)  # noqa: E731
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            call()
            torch.cuda.synchronize()
            ms.append(solver._engine.last_kernel_ms()[0])
        info = solver._engine.kernel_info()
        print('%-24s %-8s kernel %.3f ms  %.3e solves/s  regs %d  failed %d' % (
            defs or 'default (lane groups)', 'fwd+sens' if sens else 'forward', np.mean(ms), B / (np.mean(ms) * 1e-3),
            info['regs_fwd'], int((st != 0).sum())), flush=True)
        del solver
