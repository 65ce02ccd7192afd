#!/usr/bin/env python
"""Benchmark of the hot path: batched IVP solves (forward BDF + adjoint gradient) per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lv_adj] [--impl reference]

A *step* is one pass of the hot path over one batch of synthetic parameter draws: for the default
workload ``lv_adj`` (BASELINE.json configs[2], the configuration the metric "fwd+adjoint
solves/s" is quoted on) that is 65 536 Lotka-Volterra forward+adjoint solves at rtol=atol=1e-8
per GPU.  Multi-GPU runs (launched by torchrun, one rank per GPU) are weak-scaling: every rank
solves its own 65 536 draws; the only collective is the all-gather of the outputs.

One JSON line is printed by rank 0; see DESIGN.md "Measurement" for every key.  The headline
(`value`, `e2e`, `roofline`, `cpu_baseline`, `clocks`) is ``lv_adj`` on the reference's backward
schedule; the other BASELINE.json configs -- ``lv_fwd`` (configs[1]), ``robertson_adj``
(configs[3]), ``seir_adj`` at 32 768 draws per GPU (at N = 8 that is configs[4]) -- and the opt-in
restart-free backward pass are measured right after it by the same invocation, each with its own
`roofline`, `cpu_baseline`, `e2e` and `clocks`, and appended as `secondary` entries.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'ivp_solves_per_sec_fwd_adjoint'
L2_FLUSH_BYTES = 512 << 20


BACKWARD_TOL = 1e-10        # --backward-tol
INTERPOLATION = 'polynomial'   # --interpolation
BACKWARD = 'reference'         # --backward


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='lv_adj')
    ap.add_argument('--batch', type=int, default=None, help='instances per GPU (default: the workload\'s)')
    ap.add_argument('--block', type=int, default=None)
    ap.add_argument('--min-blocks', type=int, default=None)
    ap.add_argument('--no-gather', action='store_true', help='skip the output all-gather at N > 1')
    ap.add_argument('--cpu-sample', type=int, default=None, help='instances in the CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-secondary', action='store_true',
                    help='only the headline workload (default: the other BASELINE.json configs and the '
                         'restart-free backward pass follow as `secondary` entries of the same line)')
    ap.add_argument('--interpolation', default='polynomial', choices=['polynomial', 'hermite'],
                    help='AdjointSolver(interpolation=...) (reference default: polynomial)')
    ap.add_argument('--backward', default='reference', choices=['reference', 'fundamental'],
                    help='backward schedule: the reference\'s restart per output time (default, the '
                         'parity path) or the restart-free fundamental-matrix pass (opt-in, '
                         'csrc/sb_fund.cuh; the CPU arms always run the reference schedule)')
    ap.add_argument('--backward-tol', type=float, default=1e-10,
                    help='rtol = atol of the backward problem and its quadrature (the reference '
                         'hard-codes 1e-10, solver.py:599,614; README.md:243-249 shows the override)')
    return ap.parse_args()


def algorithmic_bytes_per_solve(n_s, n_all, n_deriv, n_t, adjoint, mean_fwd_steps, sens=False):
    """I/O of one solve plus, for the adjoint, every accepted forward step's
    (t, order, y) written once and read once; with forward sensitivities the initial and the
    output sensitivities."""
    b = 8 * (n_s + n_all) + 8 * n_t * n_s
    if sens:
        b += 8 * n_deriv * n_s * (n_t + 1)
    if adjoint:
        b += 8 * n_t * n_s + 8 * (n_deriv + n_s) + 2 * 8 * mean_fwd_steps * (n_s + 2)
    return float(b)


class ClockSampler:
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.err = [], set(), None, None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception as err:  # noqa: BLE001
            self.err = repr(err)
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
            'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4),
            'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
            'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
            'hw_power_brake_slowdown': getattr(nv, 'nvmlClocksEventReasonHwPowerBrakeSlowdown', 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for name, bit in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception as err:  # noqa: BLE001
                self.err = repr(err)
                break
            self._stop.wait(0.01)

    def start(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        out = {'sm_mhz': float(np.median(self.samples)) if self.samples else None,
               'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
               'samples': len(self.samples)}
        if self.err:
            out['error'] = self.err
        return out


def host_threads():
    """All host threads this process may use (torchrun pins OMP_NUM_THREADS=1 per rank, which
    must not throttle the CPU arm)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(w, problem, n_sample, adjoint, threads=0, repeats=1):
    """The oracle (CPU restatement of the reference path) on the host cores, bounded sample."""
    from oracle.oracle import Oracle, max_threads
    orc = Oracle(problem, rtol=1e-8, atol=1e-8, rtol_b=BACKWARD_TOL, atol_b=BACKWARD_TOL,
                 rtol_q=BACKWARD_TOL, atol_q=BACKWARD_TOL, interpolation=INTERPOLATION)
    y0, theta = w.draws(n_sample)
    grads = w.grads(problem.n_states)
    cores = threads or host_threads()
    best = None
    for _ in range(repeats):
        t = time.perf_counter()
        if adjoint:
            orc.solve_adjoint(w.t0, w.tvals, y0, theta, grads, n_threads=cores)
        elif w.sens:
            orc.solve_forward_sens(w.t0, w.tvals, y0, theta, np.zeros((problem.n_params, problem.n_states)),
                                   n_threads=cores)
        else:
            orc.solve_forward(w.t0, w.tvals, y0, theta, n_threads=cores)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return n_sample / best, cores, best


def run_reference(args, w, problem, rank, world):
    if rank != 0:
        return
    n_sample = args.cpu_sample or {'lv_adj': 16384, 'lv_fwd': 65536}.get(w.name, 1024)
    from oracle.oracle import Oracle, max_threads
    orc = Oracle(problem, rtol=1e-8, atol=1e-8, rtol_b=BACKWARD_TOL, atol_b=BACKWARD_TOL,
                 rtol_q=BACKWARD_TOL, atol_q=BACKWARD_TOL, interpolation=INTERPOLATION)
    y0, theta = w.draws(n_sample)
    grads = w.grads(problem.n_states)
    cores = host_threads()

    def step():
        if w.adjoint:
            orc.solve_adjoint(w.t0, w.tvals, y0, theta, grads, n_threads=cores)
        else:
            orc.solve_forward(w.t0, w.tvals, y0, theta, n_threads=cores)

    for _ in range(args.warmup):
        step()
    t = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t
    value = n_sample * args.steps / dt
    sample = ('%d of the workload\'s %d draws per step, OpenMP over instances on all host threads'
              % (n_sample, w.batch))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'solves/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': config_dict(w, problem, n_sample, 1, extra={
            'note': 'CPU restatement of the reference path (oracle/cvodes_port.c); the reference '
                    'itself needs SUNDIALS, which is absent from this image.  Rank 0 alone runs it: '
                    'batch_per_gpu / global_batch are the CPU sample per step, whatever --gpus says'}),
        'cpu_baseline': {'value': value, 'unit': 'solves/s', 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'solves/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(w, problem, batch, n_gpus, extra=None):
    cfg = {
        'workload': w.name,
        'problem': {'n_states': problem.n_states, 'n_params': problem.n_params_total,
                    'n_deriv': problem.n_params, 'n_tvals': int(len(w.tvals))},
        'batch_per_gpu': int(batch), 'global_batch': int(batch) * n_gpus,
        'rtol': 1e-8, 'atol': 1e-8, 'rtol_backward': BACKWARD_TOL, 'atol_backward': BACKWARD_TOL,
        'method': 'BDF(1-5) + Newton/dense LU; adjoint: backward BDF restarted at every tval + quadrature',
        'interpolation': INTERPOLATION, 'backward_schedule': BACKWARD,
        'cotangent': ('ones((n_t, n_s))' if w.cotangent == 'ones' else 'seeded N(0,1) [n_t, n_s]')
                     + ' shared by all instances',
        'theta': 'theta_med * exp(%g * N(0,1)), seed %d' % (w.sigma, w.seed),
        'parallelism': 'dp%d (independent draws, contiguous shards)' % n_gpus,
    }
    if extra:
        cfg.update(extra)
    return cfg


# (workload, draws per GPU, backward schedule) measured after the headline
SECONDARY = [('lv_fwd', None, 'reference'), ('robertson_adj', None, 'reference'),
             ('seir_adj', 32768, 'reference'), ('lv_adj', None, 'fundamental'),
             ('lv_fsa', None, 'reference')]


def measure(args, name, *, steps, warmup, backward, batch, rank, world, local_rank,
            with_cpu=True, with_e2e=True):
    """One workload on this rank's GPU: the device-resident leg (`value`), the end-to-end leg
    through the public API with host buffers (`e2e`), the roofline of the dominant kernel and the
    CPU baseline.  All ranks call this together; the returned line is complete on rank 0."""
    global BACKWARD
    import torch
    import torch.distributed as dist
    from sunode_b200 import examples, sharding
    from sunode_b200._engine import PinnedBuffer
    from sunode_b200.solver import AdjointSolver, Solver

    BACKWARD = backward
    dev = torch.device('cuda', local_rank)
    w = examples.workloads()[name]
    t_setup = time.perf_counter()
    problem = w.make_problem()
    _ = problem.generated              # sympy derivation + code generation (outside every timed region)
    codegen_s = time.perf_counter() - t_setup

    B = batch or w.batch
    n_t, n_s, n_all, n_d = len(w.tvals), problem.n_states, problem.n_params_total, problem.n_params
    y0_h, theta_h = w.draws(B, offset=rank * B)
    grads_h = w.grads(n_s)
    t_setup = time.perf_counter()
    if w.adjoint:
        solver = AdjointSolver(problem, abstol=1e-8, reltol=1e-8, interpolation=INTERPOLATION,
                               backward=BACKWARD, history_capacity=w.history_capacity, device=local_rank,
                               block_threads=args.block, min_blocks=args.min_blocks)
        if BACKWARD_TOL != 1e-10:
            solver.set_backward_tolerances(BACKWARD_TOL, BACKWARD_TOL)
            solver.set_quad_tolerances(BACKWARD_TOL, BACKWARD_TOL)
    else:
        solver = Solver(problem, abstol=1e-8, reltol=1e-8, device=local_rank,
                        sens_mode='simultaneous' if w.sens else None,
                        block_threads=args.block, min_blocks=args.min_blocks)
    eng = solver._engine
    # JIT / codegen are excluded from the metric and reported separately
    setup = {'codegen_s': round(codegen_s, 3),
             'kernel_build_s': round(time.perf_counter() - t_setup, 3),
             'note': 'sympy -> CUDA source; NVRTC compile or in-tree cubin cache load + handle creation'}

    # ---- device-resident inputs / outputs (the `value` leg)
    y0_d = torch.from_numpy(y0_h).to(dev)
    theta_d = torch.from_numpy(theta_h).to(dev)
    grads_d = torch.from_numpy(grads_h).to(dev)
    y_d = torch.empty((B, n_t, n_s), dtype=torch.float64, device=dev)
    g_d = torch.empty((B, n_d), dtype=torch.float64, device=dev)
    l_d = torch.empty((B, n_s), dtype=torch.float64, device=dev)
    st_d = torch.empty((B,), dtype=torch.int32, device=dev)
    sf_d = torch.zeros((B, 8), dtype=torch.int32, device=dev)
    sb_d = torch.zeros((B, 8), dtype=torch.int32, device=dev)
    if w.sens:
        s0_d = torch.zeros((n_d, n_s), dtype=torch.float64, device=dev)      # dy0/dp = 0: y0 is fixed
        s_d = torch.empty((B, n_t, n_d, n_s), dtype=torch.float64, device=dev)
    gather = world > 1 and not args.no_gather
    counts = [B] * world
    if gather:
        if w.adjoint:
            # the forward kernel writes into symmetric memory, peers pull it with their copy engines
            sym, _ = sharding.symmetric_rows((B, n_t, n_s), dev)
            if sym is not None:
                y_d = sym
        y_all = torch.empty((world * B, n_t, n_s), dtype=torch.float64, device=dev)
        small_all = torch.empty((world * B, n_d + n_s + 1), dtype=torch.float64, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def device_step(y0, theta, grads, stats=False, host_out=None):
        """The hot path on device-resident arrays.  N > 1: through sunode_b200.sharding -- the
        trajectories' all-gather is in flight underneath the backward kernels."""
        if w.adjoint and gather and not stats:
            sharding.solve_adjoint_gathered(solver, w.t0, w.tvals, y0, theta, grads, counts,
                                            y_out=y_d, grad_out=g_d, lamda_out=l_d, status=st_d,
                                            y_all=y_all, small_all=small_all, host_out=host_out)
        elif w.adjoint:
            solver.solve_adjoint_batch(w.t0, w.tvals, y0, theta, grads, y_out=y_d,
                                       grad_out=g_d, lamda_out=l_d, status=st_d,
                                       stats_fwd=sf_d if stats else None,
                                       stats_bwd=sb_d if stats else None)
        else:
            if w.sens:
                solver.solve_sens_batch(w.t0, w.tvals, y0, theta, s0_d, y_out=y_d, sens_out=s_d,
                                        status=st_d, stats=sf_d if stats else None)
            else:
                solver.solve_batch(w.t0, w.tvals, y0, theta, y_out=y_d, status=st_d,
                                   stats=sf_d if stats else None)
            if gather:
                dist.all_gather_into_tensor(y_all, y_d)
            if host_out is not None:
                host_out['y'].copy_(y_d, non_blocking=True)
                if w.sens:
                    host_out['s'].copy_(s_d, non_blocking=True)
                host_out['st'].copy_(st_d, non_blocking=True)

    for i in range(max(warmup, 1)):
        device_step(y0_d, theta_d, grads_d, stats=(i == 0))
    torch.cuda.synchronize()
    n_fail = int((st_d != 0).sum().item())
    codes, cnts = torch.unique(st_d[st_d != 0], return_counts=True)
    fail_codes = {int(c): int(n) for c, n in zip(codes.tolist(), cnts.tolist())}
    mean_fwd_steps = float(sf_d[:, 0].double().mean().item())
    mean_bwd_steps = float(sb_d[:, 0].double().mean().item()) if w.adjoint else 0.0

    sampler = ClockSampler(local_rank)
    launches0 = eng.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    kern_ms = np.zeros((steps, 3))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    for k in range(steps):
        flush.zero_()                      # evict L2 between timed iterations (untimed)
        starts[k].record()
        device_step(y0_d, theta_d, grads_d)
        ends[k].record()
        ends[k].synchronize()
        kern_ms[k] = eng.last_kernel_ms()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    total_ms = float(sum(s.elapsed_time(e) for s, e in zip(starts, ends)))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        fails = torch.tensor([n_fail], dtype=torch.int64, device=dev)
        dist.all_reduce(fails)
        n_fail = int(fails.item())
    value = world * B * steps / (total_ms * 1e-3)

    # ---- end to end through the public API with host buffers (the `e2e` leg)
    e2e = None
    if with_e2e:
        h2d = 8 * (B * n_s + B * n_all + n_t) + (8 * n_t * n_s if w.adjoint else 0) + 8 * n_s
        d2h = 8 * B * n_t * n_s + 4 * B + (8 * B * (n_d + n_s) if w.adjoint else 0)
        shapes = {'y0': (B, n_s), 'theta': (B, n_all), 'grads': (n_t, n_s), 'y': (B, n_t, n_s),
                  'g': (B, n_d), 'l': (B, n_s)}
        if w.sens:
            shapes['s'] = (B, n_t, n_d, n_s)
            s0_h = np.zeros((n_d, n_s))
            h2d += 8 * n_d * n_s
            d2h += 8 * B * n_t * n_d * n_s
        pin = {k: PinnedBuffer(s) for k, s in shapes.items()}
        pin_st = PinnedBuffer((B,), np.int32)
        pin['y0'].array[...] = y0_h
        pin['theta'].array[...] = theta_h
        pin['grads'].array[...] = grads_h
        page = {k: np.array(v.array) for k, v in pin.items()}          # ordinary (pageable) numpy arrays
        page_st = np.zeros((B,), np.int32)

        def host_step(buf, st):
            """One call of the public API on HOST arrays: the library uploads the inputs, runs the
            kernels and downloads the results before it returns."""
            if w.adjoint:
                solver.solve_adjoint_batch(w.t0, w.tvals, buf['y0'], buf['theta'], buf['grads'],
                                           y_out=buf['y'], grad_out=buf['g'], lamda_out=buf['l'],
                                           status=st)
            elif w.sens:
                solver.solve_sens_batch(w.t0, w.tvals, buf['y0'], buf['theta'], s0_h, y_out=buf['y'],
                                        sens_out=buf['s'], status=st)
            else:
                solver.solve_batch(w.t0, w.tvals, buf['y0'], buf['theta'], y_out=buf['y'], status=st)

        if world > 1 and gather:
            # N > 1: host shard -> device, sharded solve with both all-gathers (sunode_b200.sharding),
            # this rank's results -> host; the gathered copies stay on the device
            tp = {k: torch.from_numpy(v.array) for k, v in pin.items()}
            tp_st = torch.from_numpy(pin_st.array)
            in_d = {k: torch.empty_like(tp[k], device=dev) for k in ('y0', 'theta', 'grads')}
            host = {'y': tp['y'], 'g': tp['g'], 'l': tp['l'], 'st': tp_st, 's': tp.get('s')}

            def e2e_step():
                for k in ('y0', 'theta', 'grads'):
                    in_d[k].copy_(tp[k], non_blocking=True)
                device_step(in_d['y0'], in_d['theta'], in_d['grads'], host_out=host)
                torch.cuda.current_stream().synchronize()
            api = ('pinned host shard -> device (torch copies), sunode_b200.sharding.solve_adjoint_gathered '
                   '(all-gathers included), this rank\'s results -> pinned host')
        else:
            arrays = {k: v.array for k, v in pin.items()}

            def e2e_step():
                host_step(arrays, pin_st.array)
            api = 'AdjointSolver.solve_adjoint_batch / Solver.solve_batch on pinned host arrays (C ABI, SB_MEM_HOST)'

        def timed(fn):
            for _ in range(max(warmup, 1)):
                fn()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(steps):
                fn()                       # returns after the D2H copies completed
            secs = time.perf_counter() - t
            if world > 1:
                tt = torch.tensor([secs], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                secs = float(tt.item())
            return secs

        e2e_s = timed(e2e_step)
        assert np.array_equal(pin['y'].array, y_d.cpu().numpy()), 'e2e and device legs disagree'
        e2e = {'value': world * B * steps / e2e_s, 'unit': 'solves/s',
               'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
               'ms_per_step': 1e3 * e2e_s / steps, 'api': api,
               'timer': 'host wall clock around the public API call (includes H2D, kernels, D2H'
                        + (', all-gathers)' if world > 1 and gather else ')')}
        if world == 1:
            # the same call on ordinary numpy arrays (what a user who never heard of pinned memory
            # passes): the driver stages pageable copies through its own bounce buffers
            page_s = timed(lambda: host_step(page, page_st))
            assert np.array_equal(page['y'], pin['y'].array)
            e2e['pageable'] = {'value': B * steps / page_s, 'ms_per_step': 1e3 * page_s / steps,
                               'note': 'same call, pageable numpy arrays instead of pinned ones'}
        del pin, pin_st

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel
        peaks = {}
        if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')):
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
                peaks = json.load(fh)
        peak = float(peaks.get('hbm_gbs', 6650.0))
        peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650'
        bytes_solve = algorithmic_bytes_per_solve(n_s, n_all, n_d, n_t, w.adjoint, mean_fwd_steps, w.sens)
        dom = 2 if w.adjoint else 0
        dom_ms = float(kern_ms[:, dom].mean())
        achieved = bytes_solve * B / (dom_ms * 1e-3) / 1e9
        info = eng.kernel_info()
        from sunode_b200._engine import lanes_per_instance, FLAT_FWD_STEPS_PER_TVAL
        group = lanes_per_instance(n_s)
        # which build of the backward kernel did the work (the device-side rule of sb_api.cpp)
        flat = w.adjoint and group == 1 and mean_fwd_steps * (1 - n_fail / max(B, 1)) > FLAT_FWD_STEPS_PER_TVAL * n_t
        kernel = ('sb_backward_fund' if BACKWARD == 'fundamental' else
                  'sb_backward_flat' if flat else 'sb_backward') if w.adjoint else (
                      'sb_forward_sens' if w.sens else 'sb_forward')
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get('%s:%d:%s' % (w.name, B, kernel))
        roofline = {
            'bound': 'hbm', 'kernel': kernel,
            'lanes_per_instance': {'forward': 1, 'backward': group if w.adjoint else None},
            'backward_schedule': None if not w.adjoint else (
                'restart-free (fundamental matrix), one lane per instance' if BACKWARD == 'fundamental' else
                'every lane walks its intervals on its own' if flat else 'lanes of a warp restart together'),
            'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
            'traffic': traffic, 'traffic_source': 'ncu --set full capture, profiles/ncu_traffic.json' if traffic else None,
            'peak_source': peak_src,
            'algorithmic_bytes_per_solve': bytes_solve, 'kernel_ms': dom_ms,
            'kernel_ms_all': {'sb_forward': float(kern_ms[:, 0].mean()),
                              'sb_tables': float(kern_ms[:, 1].mean()),
                              'sb_backward': float(kern_ms[:, 2].mean())},
            'kernel_share_of_step': dom_ms / (total_ms / steps),
            'note': 'the path is FP64-latency bound, not HBM bound (DESIGN.md); frac is reported as '
                    'the contract asks, warp efficiency and FP64 pipe use are in profiles/',
            'mean_steps': {'forward': mean_fwd_steps, 'backward': mean_bwd_steps},
            'registers': {'forward': info['regs_fwd'], 'backward': info['regs_bwd']},
            'block_threads': info['block_threads'],
            'resident_blocks_per_sm': {'forward': info['blocks_per_sm_fwd'],
                                       'backward': info['blocks_per_sm_bwd']},
        }
        line = {
            'metric': METRIC, 'value': value, 'unit': 'solves/s', 'n_gpus': world,
            'steps': steps, 'warmup': warmup, 'ms_per_step': total_ms / steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': config_dict(w, problem, B, world, extra={
                'l2': 'L2 flushed between timed iterations (512 MiB memset, untimed)',
                'failed_instances': n_fail, 'failed_status_rank0': fail_codes,
                'collective': ('all_gather(y_out) issued after the forward kernel, in flight under the '
                               'backward kernels (%s), + NCCL all_gather(grad|lamda0|status) per step'
                               % ('peer-to-peer pulls from symmetric memory by the copy engines'
                                  if sharding.last_transport() == 'p2p' else 'NCCL')
                               if gather and w.adjoint else 'all_gather(y_out) per step' if gather else 'none')}),
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline,
            'setup': setup,
        }
        if with_cpu and world == 1:
            # a bounded sample: ~20 s of CPU work for the LV workload, spread over the host threads
            n_sample = args.cpu_sample or min(B, {'lv_adj': 32768, 'lv_fwd': 65536}.get(w.name, 2048))
            v, cores, secs = cpu_baseline(w, problem, n_sample, w.adjoint)
            line['cpu_baseline'] = {
                'value': v, 'unit': 'solves/s', 'cores': cores, 'kind': 'port',
                'sample': 'first %d of the %d draws, OpenMP over instances, %.1f s' % (n_sample, B, secs)}
    del solver, eng, flush, y_d
    if gather:
        del y_all, small_all
    torch.cuda.empty_cache()
    return line


def main():
    global BACKWARD_TOL, INTERPOLATION, BACKWARD
    args = parse_args()
    BACKWARD_TOL = float(args.backward_tol)
    INTERPOLATION = args.interpolation
    BACKWARD = args.backward if args.impl != 'reference' else 'reference'   # CPU arms: reference schedule
    from sunode_b200 import examples
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        w = examples.workloads()[args.workload]
        run_reference(args, w, w.make_problem(), rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    common = dict(rank=rank, world=world, local_rank=local_rank)
    line = measure(args, args.workload, steps=args.steps, warmup=args.warmup, backward=BACKWARD,
                   batch=args.batch, with_cpu=not args.no_cpu_baseline, with_e2e=not args.no_e2e,
                   **common)
    headline_default = (args.workload == 'lv_adj' and BACKWARD == 'reference' and args.batch is None
                        and INTERPOLATION == 'polynomial' and BACKWARD_TOL == 1e-10)
    if headline_default and not args.no_secondary:
        secondary = []
        for name, batch, backward in SECONDARY:
            entry = measure(args, name, steps=max(2, min(args.steps, 5)), warmup=3, backward=backward,
                            batch=batch, with_cpu=not args.no_cpu_baseline, with_e2e=not args.no_e2e,
                            **common)
            if entry is not None:
                secondary.append(entry)
        if line is not None:
            line['secondary'] = secondary
    if line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
