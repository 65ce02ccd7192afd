"""Where do the idle lanes of sb_backward come from?  (CPU-side analysis with the host emulation.)

Walks the backward pass of a sample of draws interval by interval (the emulated device code,
tests/emu) and reads the per-interval pass counts (accepted steps + failed error tests + Newton
failures) of every draw.  Lanes of a warp walk the intervals together, so a warp spends
max-over-lanes passes in an interval: the printed ratio mean/max is the lane utilisation that the
interval barrier alone allows (divergence inside a pass comes on top).
"""
import ctypes, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sunode_b200 import examples
from tests.emu import emu as E

name = sys.argv[1] if len(sys.argv) > 1 else 'lv_adj'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
w = examples.workloads()[name]
prob = w.make_problem()
em = E.Emulator(prob, os.path.join(tempfile.gettempdir(), 'sb_emu_tools'))
y0, theta = w.draws(B)
g = w.grads(em.ns)
cap = w.history_capacity
tab = np.zeros((B, cap, 10 + 6 * em.ns))
fwd = em.forward(w.t0, w.tvals, y0, theta, 1e-8, 1e-8, hist_cap=cap, max_steps=2 ** 30, tab=tab)
n_t = len(w.tvals)
ndq = max(em.nd, 1)
grad_out = np.zeros((B, ndq)); lam_out = np.zeros((B, em.ns))
status = np.zeros(B, np.int32); stats = np.zeros((B, 8), np.int32)
carry_d = np.zeros((B, em.ns + ndq)); carry_i = np.zeros((B, 10), np.int32)
ba = E.BackwardArgs(1e-10, 1e-10, 1e-10, 1e-10, float(w.tvals[-1]), float(w.t0),
                    E._dp(fwd['tvals']), E._dp(fwd['params']), E._dp(g), E._dp(tab),
                    E._ip(fwd['hist_n']), E._ip(fwd['status']), E._dp(grad_out), E._dp(lam_out),
                    E._ip(status), E._ip(stats), B, n_t, cap, 25000, 1, None, None,
                    None, None, E._dp(carry_d), E._ip(carry_i), n_t + 1, 1, 0, 32, -1, 0, None, 0)
em.lib.emu_backward_unit.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
passes = np.zeros((n_t + 1, B), np.int64)
prev = np.zeros(B, np.int64)
for k in range(n_t + 1):
    em.lib.emu_backward_unit(ctypes.byref(ba), k, k + 1)
    if k < n_t:
        cur = carry_i[:, 2].astype(np.int64) + carry_i[:, 6] + carry_i[:, 7]
    else:
        cur = stats[:, 0].astype(np.int64) + stats[:, 4] + stats[:, 5]
    passes[k] = cur - prev
    prev = cur
print('workload %s, %d draws: mean passes/solve %.1f, failures %d' % (name, B, passes.sum(0).mean(), (status != 0).sum()))
for lanes in (32, 16, 8, 4):
    p = passes[:, :B // lanes * lanes].reshape(n_t + 1, -1, lanes)
    per_interval = p.sum() / (lanes * p.max(axis=2).sum())
    tot = p.sum(axis=0)                                   # no barrier: only the end of the solve
    whole = tot.sum() / (lanes * tot.max(axis=1).sum())
    print('  %2d lanes/warp: utilisation with a barrier per interval %.3f, without %.3f'
          % (lanes, per_interval, whole))

# ---- restart policies of the flattened backward kernel ------------------------------------------
# A lane that has finished its interval waits; all waiting lanes restart together (one execution of
# the divergent restart block, costing R pass-equivalents of warp time) as soon as `need` lanes
# wait, or the oldest waiter has waited `patience` passes, or nobody is left stepping.
def simulate(p, need, patience, R=0.3):
    n_int, n_warp, lanes = p.shape
    total = 0.0
    for w in range(n_warp):
        rem = p[0, w].copy(); k = np.zeros(lanes, int); waited = np.zeros(lanes, int)
        # intervals with zero passes are skipped without a restart
        t = 0.0
        while True:
            waiting = (rem == 0) & (k < n_int)
            stepping = rem > 0
            if not waiting.any() and not stepping.any():
                break
            if waiting.any() and (waiting.sum() >= need or waited[waiting].max() >= patience or not stepping.any()):
                idx = np.nonzero(waiting)[0]
                k[idx] += 1
                live = idx[k[idx] < n_int]
                rem[live] = p[k[live], w, live]
                waited[idx] = 0
                if (rem[live] > 0).any():
                    t += R
                continue
            # advance to the next event: the smallest remaining count among stepping lanes
            adv = rem[stepping].min()
            if waiting.any():
                adv = min(adv, max(1, patience - waited[waiting].max()))
            rem[stepping] -= adv
            waited[waiting] += adv
            t += adv
        total += t
    return p.sum() / (lanes * total)

if len(sys.argv) > 3:
    p32 = passes[:, :B // 32 * 32].reshape(n_t + 1, -1, 32)[:, :int(sys.argv[3])]
    for need, patience in ((32, 10 ** 9), (1, 0), (4, 4), (8, 8), (8, 16), (16, 16), (16, 32), (24, 64)):
        print('  policy need=%2d patience=%-10d utilisation %.3f' % (need, patience, simulate(p32, need, patience)))
