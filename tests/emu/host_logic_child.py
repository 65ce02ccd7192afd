"""Child process of tests/test_host_logic.py: runs the product's Python layer and C-ABI library on
top of tests/emu/fake_cuda.cpp (LD_LIBRARY_PATH is set by the parent) and compares what comes out
with the same emulated kernels called directly.  TEST INFRASTRUCTURE ONLY."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from sunode_b200 import _lib, examples                       # noqa: E402
from sunode_b200.solver import AdjointSolver, Solver         # noqa: E402
from tests.emu.emu import Emulator                           # noqa: E402
from tests.test_options import chase_inputs, chase_problem, constraint_define   # noqa: E402

work = tempfile.mkdtemp()
assert _lib.device_count() == 1

# the fake driver never looks inside a cubin (the kernels it runs are the emulation library's), so
# NVRTC -- which would talk to the real driver -- is kept out of this process
from sunode_b200 import _engine                               # noqa: E402
_engine.compile_cubin = lambda gen, **kw: (b'\x7fELF' + bytes(60), '<not compiled>')


def emulator(problem, defines=()):
    emu = Emulator(problem, work, defines=defines)
    os.environ['SB_FAKE_EMU_LIB'] = emu.lib._name      # read by the fake driver at module load
    return emu


w = examples.workloads()['lv_adj']
prob = w.make_problem()
B = 40                                                   # not a multiple of 32: padding lanes
y0, theta = w.draws(B)
grads = np.random.default_rng(1).standard_normal((B, len(w.tvals), 2))

# 1. the default path: fused call, two-call form, shared cotangent, stats routing
emu = emulator(prob)
ref = emu.adjoint(w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=512)
solver = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512)
sf, sb = np.zeros((B, 8), np.int32), np.zeros((B, 8), np.int32)
y, g, lam, st = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads, stats_fwd=sf, stats_bwd=sb)
assert (st == 0).all()
for a, b in ((y, ref['y']), (g, ref['grad']), (lam, ref['lamda']), (sf, ref['fwd']['stats']), (sb[:, :7], ref['stats'][:, :7])):
    np.testing.assert_array_equal(a, b)
y2, st2 = solver.solve_forward_batch(w.t0, w.tvals, y0, theta)
g2, l2, sb2 = solver.solve_backward_batch(w.tvals[-1], w.t0, w.tvals, grads)
np.testing.assert_array_equal(g2, g)
np.testing.assert_array_equal(l2, lam)
# the same call with mem = SB_MEM_DEVICE (what torch CUDA tensors select; here "device" memory is
# host memory): no staging, status copied device-to-device, asynchronous until sb_synchronize
import ctypes                                                # noqa: E402
eng = solver._engine
yd, gd, ld = np.full_like(y, -1.0), np.full_like(g, -1.0), np.full_like(lam, -1.0)
std = np.full(B, 77, np.int32)
ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)            # noqa: E731
tv = np.ascontiguousarray(w.tvals)
_lib.check(eng._lib.sb_solve_adjoint(eng._h, B, float(w.t0), tv.ctypes.data, len(tv), ptr(y0), ptr(theta),
                                     ptr(grads), 0, ptr(yd), ptr(gd), ptr(ld), ptr(std), None, None,
                                     _lib.SB_MEM_DEVICE, None))
eng.synchronize()
for a, b in ((yd, y), (gd, g), (ld, lam), (std, st)):
    np.testing.assert_array_equal(a, b)
assert eng.launch_count() > 0 and all(ms >= 0 for ms in eng.last_kernel_ms())
print('default path ok')

# 2. Hermite: the history stride comes back from the module, tables are cubic entries
emu = emulator(prob, ('SB_HERMITE',))
ref = emu.adjoint(w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=512)
out = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512,
                    interpolation='hermite').solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads)
assert (out[3] == 0).all()
np.testing.assert_array_equal(out[1], ref['grad'])
np.testing.assert_array_equal(out[2], ref['lamda'])
assert not np.array_equal(out[1], g)
print('hermite ok')

# 3. restart-free backward pass: kernel selection in the launcher
emu = emulator(prob, ('SB_FUND',))
ref = emu.adjoint(w.t0, w.tvals, y0, theta, grads, 1e-8, 1e-8, hist_cap=512, fund=True)
fs = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512, backward='fundamental')
sb = np.zeros((B, 8), np.int32)
out = fs.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads, stats_bwd=sb)
assert (out[3] == 0).all()
np.testing.assert_array_equal(out[1], ref['grad'])
np.testing.assert_array_equal(out[2], ref['lamda'])
np.testing.assert_array_equal(sb, ref['stats'])
assert sb[:, 0].mean() < 0.25 * ref['fwd']['stats'][:, 0].mean() * 20      # ~250 against ~1 660
np.testing.assert_allclose(out[1], g, rtol=0, atol=1e-7 * np.abs(g).max())
# traces of the restart-free pass against the reference schedule's (separate solver, same driver)
fs.solve_forward_batch(w.t0, w.tvals, y0, theta)
la, qa = np.empty((B, 50, 2)), np.empty((B, 50, 2))
fs.solve_backward_batch(w.tvals[-1], w.t0, w.tvals, grads, lamda_all_out=la, quad_all_out=qa)
emulator(prob)
rs = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512)
rs.solve_forward_batch(w.t0, w.tvals, y0, theta)
lb, qb = np.empty_like(la), np.empty_like(qa)
rs.solve_backward_batch(w.tvals[-1], w.t0, w.tvals, grads, lamda_all_out=lb, quad_all_out=qb)
assert np.isfinite(la).all() and np.isfinite(qa).all()
assert np.max(np.abs(la - lb)) <= 1e-6 * np.abs(lb).max(), np.max(np.abs(la - lb))
assert np.max(np.abs(qa - qb)) <= 1e-6 * np.abs(qb).max(), np.max(np.abs(qa - qb))
print('fundamental ok')

# 4. history growth: default capacity forced down, the solve is repeated with a larger store
emulator(prob)
auto = AdjointSolver(prob, abstol=1e-8, reltol=1e-8)
auto._history_capacity = 32
auto._engine.set_history_capacity(32)
out = auto.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads)
assert auto._history_capacity in (128, 512) and (out[3] == 0).all(), auto._history_capacity   # 32 -> 128 (-> 512)
np.testing.assert_array_equal(out[1], g)
fixed = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=32)
st = fixed.solve_forward_batch(w.t0, w.tvals, y0, theta)[1]
assert (st == -1).all()
print('history growth ok')

# 5. forward sensitivities with scaling factors: per-component tolerance array
pbar = np.array([1e4, -1e3])
emu = emulator(prob)
s0 = np.zeros((2, 2))
ref = emu.forward_sens(w.t0, w.tvals, y0, theta, s0, 1e-6, 1e-6, pbar=pbar)
ys, ss, st = Solver(prob, abstol=1e-6, reltol=1e-6, sens_mode='simultaneous',
                    scaling_factors=pbar).solve_sens_batch(w.t0, w.tvals, y0, theta, s0)
assert (st == 0).all()
np.testing.assert_array_equal(ss, ref['sens'])
ref1 = emu.forward_sens(w.t0, w.tvals, y0, theta, s0, 1e-6, 1e-6)
ss1 = Solver(prob, abstol=1e-6, reltol=1e-6, sens_mode='simultaneous').solve_sens_batch(
    w.t0, w.tvals, y0, theta, s0)[1]
np.testing.assert_array_equal(ss1, ref1['sens'])
assert not np.array_equal(ss1, ss)
print('sens scaling ok')

# 6. constraints: status codes and NaN rows reach the caller
cprob = chase_problem()
cy0, ctheta, ctv = chase_inputs(16)
for cons in ([0.0, 1.0], [2.0, 1.0]):
    emu = emulator(cprob, (constraint_define(cons),))
    ref = emu.forward(0.0, ctv, cy0, ctheta, 1e-4, 1e-7)
    yc, stc = Solver(cprob, abstol=1e-7, reltol=1e-4, constraints=np.array(cons)).solve_batch(
        0.0, ctv, np.tile(cy0, (16, 1)), ctheta)
    np.testing.assert_array_equal(stc, ref['status'])
    np.testing.assert_array_equal(yc, ref['y'])
print('constraints ok')

# 7. bounded memory: a workspace limit below the batch's store cuts sb_solve_adjoint into chunks
#    (forward + backward chunk by chunk); results are those of the single launch, bit for bit
emulator(prob)
cs = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512)
per_instance = 512 * ((2 + 2) + (10 + 6 * 2)) * 8
cs.set_workspace_limit(9 * per_instance)                 # 40 draws -> chunks of 9
la, qa = np.empty((B, 50, 2)), np.empty((B, 50, 2))
sfc, sbc = np.zeros((B, 8), np.int32), np.zeros((B, 8), np.int32)
yc, gc, lc, stc = cs.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads, stats_fwd=sfc, stats_bwd=sbc)
assert cs._engine.last_chunks() == 5, cs._engine.last_chunks()
for a, b in ((yc, y), (gc, g), (lc, lam), (stc, np.zeros(B, np.int32)), (sfc, sf)):
    np.testing.assert_array_equal(a, b)
gs = np.ones((len(w.tvals), 2))                          # shared cotangent through the chunk loop
g_one = solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, gs)[1]
np.testing.assert_array_equal(cs.solve_adjoint_batch(w.t0, w.tvals, y0, theta, gs)[1], g_one)
try:                                                     # the chunked store holds the last chunk only
    cs.solve_backward_batch(w.tvals[-1], w.t0, w.tvals, grads)
    raise AssertionError('solve_backward after a chunked solve_adjoint must be refused')
except _lib.LibraryError as err:
    assert 'stored forward solve' in str(err)
cs.set_workspace_limit(per_instance // 2)
try:
    cs.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads)
    raise AssertionError('a store that cannot hold one instance must be refused')
except _lib.LibraryError as err:
    assert 'workspace limit' in str(err)
print('chunking ok')

# 8. the reference's error text names the output time that was being integrated to
from sunode_b200 import SympyProblem                      # noqa: E402
from sunode_b200.basic import SolverError                 # noqa: E402
bprob = SympyProblem({'k': ()}, {'x': ()}, lambda t, y, p: {'x': p.k * y.x ** 2}, [('k',)])
emulator(bprob)
bs = Solver(bprob, abstol=1e-8, reltol=1e-8)
btv = np.linspace(0.1, 2, 20)
bs.set_params(np.array((1.0,), dtype=bprob.params_dtype)[()])      # x = 1 / (1 - t): blows up at t = 1
try:
    bs.solve(0.0, btv, np.ones(1), bs.make_output_buffers(btv))
    raise AssertionError('expected SolverError')
except SolverError as err:
    assert 'before time=%s' % btv[9] in str(err), str(err)      # 0.1 + 9 * 0.1, just short of the pole
print('failing time ok')

# 9. armed traces are dropped by a call that fails validation; parameters of the stored forward
#    pass are the handle's own copy
emulator(prob)
ts = AdjointSolver(prob, abstol=1e-8, reltol=1e-8, history_capacity=512)
th = theta.copy()
ts.solve_forward_batch(w.t0, w.tvals, y0, th)
th[...] = np.nan                                          # the caller's array dies / is overwritten
la = np.full((B, 50, 2), 7.0)
try:                                                      # wrong n_t: refused before any launch
    ts._engine.backward(w.tvals[-1], w.t0, w.tvals[:-1], None, grads[:, :-1], np.empty((B, 2)),
                        np.empty((B, 2)), np.empty(B, np.int32), lamda_all=la[:, :-1].copy())
    raise AssertionError('expected a state error')
except _lib.LibraryError:
    pass
g9, l9, st9 = ts.solve_backward_batch(w.tvals[-1], w.t0, w.tvals, grads)   # no trace requested
assert (st9 == 0).all() and (la == 7.0).all()
np.testing.assert_array_equal(g9, g)
np.testing.assert_array_equal(l9, lam)
print('trace / params lifetime ok')
print('ALL OK')
