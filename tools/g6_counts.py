#!/usr/bin/env python
"""Work counters of the CUDA path on SUNDIALS' cvRoberts_dns problem (see tests/test_oracle.py::test_g6)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sunode_b200 import examples
from sunode_b200.solver import Solver
prob = examples.robertson()
tv = 0.4 * 10.0 ** np.arange(12)
atol = np.array([1e-8, 1e-14, 1e-6])
B = 64
rng = np.random.default_rng(0)
y0 = np.tile([1.0, 0.0, 0.0], (B, 1))
th = np.array([0.04, 3e7, 1e4]) * (1 + 1e-3 * rng.standard_normal((B, 3)))
th[0] = [0.04, 3e7, 1e4]
stats = np.zeros((B, 8), dtype=np.int32)
y, st = Solver(prob, abstol=atol, reltol=1e-4).solve_batch(0.0, tv, y0, th, stats=stats, max_retries=10)
print(os.environ.get('SUNODE_B200_DEFINES', '-'), 'exact instance:', stats[0, :7], 'status', st[0])
print('   64 draws perturbed by 0.1 %%: nst mean %.1f min %d max %d; netf mean %.1f; nje mean %.1f' % (
    stats[:, 0].mean(), stats[:, 0].min(), stats[:, 0].max(), stats[:, 4].mean(), stats[:, 2].mean()))
