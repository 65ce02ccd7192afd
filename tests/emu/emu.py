"""Host emulation of the device code (test infrastructure, see cuda_shim.h).

Builds the generated ``__device__`` functions together with ``sb_kernels.cuh`` with g++ and runs
the per-instance integrators on the CPU, so that the CPU-only test tier can compare the *device*
integrator logic with the oracle.  Nothing in the product imports this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.normpath(os.path.join(_HERE, '..', '..', 'sunode_b200', 'csrc'))
_DP = ctypes.POINTER(ctypes.c_double)
_IP = ctypes.POINTER(ctypes.c_int)
STATS = 8


class ForwardArgs(ctypes.Structure):
    _fields_ = [('t0', ctypes.c_double), ('rtol', ctypes.c_double), ('tvals', _DP), ('y0', _DP),
                ('params', _DP), ('atol', _DP), ('y_out', _DP), ('hist', _DP), ('hist_n', _IP),
                ('status', _IP), ('stats', _IP), ('B', ctypes.c_longlong), ('n_t', ctypes.c_int),
                ('hist_cap', ctypes.c_int), ('max_steps', ctypes.c_int),
                ('sens0_shared', ctypes.c_int), ('lanes', ctypes.c_int), ('pad_', ctypes.c_int), ('sens0', _DP), ('sens_out', _DP), ('tab', _DP), ('steps_total', ctypes.c_void_p), ('fail_k', _IP)]


class TablesArgs(ctypes.Structure):
    _fields_ = [('hist', _DP), ('hist_n', _IP), ('tab', _DP), ('B', ctypes.c_longlong),
                ('hist_cap', ctypes.c_int), ('pad_', ctypes.c_int)]


class BackwardArgs(ctypes.Structure):
    _fields_ = [('rtol', ctypes.c_double), ('atol', ctypes.c_double), ('rtol_q', ctypes.c_double),
                ('atol_q', ctypes.c_double), ('t_start', ctypes.c_double),
                ('t_end', ctypes.c_double), ('tvals', _DP), ('params', _DP), ('grads', _DP),
                ('tab', _DP), ('hist_n', _IP), ('fwd_status', _IP), ('grad_out', _DP),
                ('lamda_out', _DP), ('status', _IP), ('stats', _IP), ('B', ctypes.c_longlong),
                ('n_t', ctypes.c_int), ('hist_cap', ctypes.c_int), ('max_steps', ctypes.c_int),
                ('grads_shared', ctypes.c_int), ('lamda_all', _DP), ('quad_all', _DP),
                ('queue', _IP), ('seg_done', _IP), ('carry_d', _DP), ('carry_i', _IP),
                ('n_seg', ctypes.c_int), ('seg_len', ctypes.c_int), ('n_groups', ctypes.c_int),
                ('lanes', ctypes.c_int), ('flat', ctypes.c_int), ('pad2_', ctypes.c_int),
                ('steps_total', ctypes.c_void_p), ('flat_steps', ctypes.c_ulonglong)]


def _dp(a):
    return None if a is None else a.ctypes.data_as(_DP)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_IP)


class Emulator:
    def __init__(self, problem, workdir, defines=(), group=False):
        """``group=True`` additionally builds the grouped-lane backward driver (csrc/sb_group.cuh,
        its lanes emulated by host threads, see cuda_shim_group.h); ``adjoint(..., group=True)``
        then runs the backward pass through it."""
        if group:
            defines = tuple(defines) + ('SB_HOST_EMULATION_GROUP',)
        gen = problem.generated
        self.ns, self.np, self.nd = gen.n_states, gen.n_params, gen.n_deriv
        self.hist_stride = (2 * self.ns + 2) if 'SB_HERMITE' in defines else self.ns + 2
        os.makedirs(workdir, exist_ok=True)
        inc = os.path.join(workdir, 'generated_problem.inc')
        with open(inc, 'w') as fh:
            fh.write(gen.cuda)
        out = os.path.join(workdir, 'emu_%s%s.so' % (gen.digest, ''.join('_' + d.replace('=', '') for d in defines)))
        cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
        cmd = [cxx, '-O2', '-std=c++17', '-fPIC', '-shared', '-fopenmp', '-pthread', '-ffp-contract=off',
               *['-D' + d for d in defines],
               '-I', workdir, '-I', _HERE, '-I', _CSRC, os.path.join(_HERE, 'emu_main.cpp'),
               '-o', out]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError('emulation build failed:\n' + proc.stderr)
        self.lib = ctypes.CDLL(out)

    def _prep(self, y0, params):
        y0 = np.atleast_2d(np.asarray(y0, dtype=np.float64))
        params = np.asarray(params, dtype=np.float64)
        if params.ndim == 1:
            params = params[None]
        B = max(len(y0), len(params))
        y0 = np.ascontiguousarray(np.broadcast_to(y0, (B, self.ns)))
        params = np.ascontiguousarray(np.broadcast_to(params, (B, self.np))) if self.np else np.zeros((B, 1))
        return y0, params, B

    def forward(self, t0, tvals, y0, params, rtol, atol, hist_cap=0, max_steps=2500, tab=None, group=False):
        tvals = np.ascontiguousarray(tvals, dtype=np.float64)
        y0, params, B = self._prep(y0, params)
        n_t = len(tvals)
        atol = np.ascontiguousarray(np.broadcast_to(np.asarray(atol, dtype=np.float64), (self.ns,)))
        y_out = np.zeros((B, n_t, self.ns))
        status = np.zeros(B, dtype=np.int32)
        stats = np.zeros((B, STATS), dtype=np.int32)
        hist = np.zeros((B, hist_cap, self.hist_stride)) if hist_cap else None
        hist_n = np.zeros(B, dtype=np.int32)
        a = ForwardArgs(t0, rtol, _dp(tvals), _dp(y0), _dp(params), _dp(atol), _dp(y_out),
                        _dp(hist), _ip(hist_n), _ip(status), _ip(stats), B, n_t, hist_cap,
                        max_steps, 0, 32, 0, None, None, _dp(tab), None)
        if group:
            assert self.lib.emu_group_size() > 1, 'this problem does not run in lane groups'
            self.lib.emu_forward_group(ctypes.byref(a))
        else:
            self.lib.emu_forward(ctypes.byref(a))
        return dict(y=y_out, status=status, stats=stats, hist=hist, hist_n=hist_n,
                    params=params, tvals=tvals)

    def forward_sens(self, t0, tvals, y0, params, sens0, rtol, atol, max_steps=2500, pbar=None, group=False):
        tvals = np.ascontiguousarray(tvals, dtype=np.float64)
        y0, params, B = self._prep(y0, params)
        n_t = len(tvals)
        # what sb_api.cpp uploads: atol per stacked component, sensitivity block k with atol / |pbar_k|
        atol = np.broadcast_to(np.asarray(atol, dtype=np.float64), (self.ns,))
        scale = np.ones(self.nd) if pbar is None else np.abs(np.asarray(pbar, dtype=np.float64))
        atol = np.ascontiguousarray(np.concatenate([atol] + [atol / scale[k] for k in range(self.nd)]))
        sens0 = np.ascontiguousarray(sens0, dtype=np.float64)
        shared = int(sens0.ndim == 2)
        y_out = np.zeros((B, n_t, self.ns))
        sens_out = np.zeros((B, n_t, self.nd, self.ns))
        status = np.zeros(B, dtype=np.int32)
        stats = np.zeros((B, STATS), dtype=np.int32)
        a = ForwardArgs(t0, rtol, _dp(tvals), _dp(y0), _dp(params), _dp(atol), _dp(y_out),
                        None, None, _ip(status), _ip(stats), B, n_t, 0, max_steps, shared, 32, 0,
                        _dp(sens0), _dp(sens_out), None, None)
        if group:
            assert self.lib.emu_group_size() > 1, 'this problem does not run in lane groups'
            self.lib.emu_forward_sens_group(ctypes.byref(a))
        else:
            self.lib.emu_forward_sens(ctypes.byref(a))
        return dict(y=y_out, sens=sens_out, status=status, stats=stats)

    def adjoint(self, t0, tvals, y0, params, grads, rtol, atol, rtol_b=1e-10, atol_b=1e-10,
                rtol_q=1e-10, atol_q=1e-10, hist_cap=1024, max_steps_b=25000, flat=None,
                group=False, fund=False):
        B0 = max(len(np.atleast_2d(y0)), len(np.atleast_2d(params)))
        tab_fused = np.zeros((B0, hist_cap, 10 + 6 * self.ns))
        fwd = self.forward(t0, tvals, y0, params, rtol, atol, hist_cap=hist_cap,
                           max_steps=2 ** 30, tab=tab_fused)
        B = len(fwd['status'])
        n_t = len(fwd['tvals'])
        tab = np.zeros((B, hist_cap, 10 + 6 * self.ns))
        ta = TablesArgs(_dp(fwd['hist']), _ip(fwd['hist_n']), _dp(tab), B, hist_cap, 0)
        self.lib.emu_tables(ctypes.byref(ta))
        # the forward kernel's own tables (built step by step) must be the stand-alone ones
        for b in range(B):
            n = fwd['hist_n'][b]
            assert np.array_equal(tab[b, 1:n], tab_fused[b, 1:n]), 'fused tables differ'
        grads = np.ascontiguousarray(grads, dtype=np.float64)
        shared = int(grads.ndim == 2)
        grad_out = np.zeros((B, max(self.nd, 1)))[:, :self.nd].copy() if self.nd else np.zeros((B, 0))
        grad_out = np.ascontiguousarray(grad_out)
        lam_out = np.zeros((B, self.ns))
        status = np.zeros(B, dtype=np.int32)
        stats = np.zeros((B, STATS), dtype=np.int32)
        gptr = _dp(grad_out) if self.nd else _dp(np.zeros(1))
        ba = BackwardArgs(rtol_b, atol_b, rtol_q, atol_q, float(fwd['tvals'][-1]), float(t0),
                          _dp(fwd['tvals']), _dp(fwd['params']), _dp(grads), _dp(tab),
                          _ip(fwd['hist_n']), _ip(fwd['status']), gptr, _dp(lam_out),
                          _ip(status), _ip(stats), B, n_t, hist_cap, max_steps_b, shared, None, None,
                          None, None, None, None, 1, n_t + 1, 0, 32, -1 if flat is None else flat, 0, None, 0)
        if fund:
            self.lib.emu_backward_fund(ctypes.byref(ba))
        elif group:
            assert self.lib.emu_group_size() > 1, 'this problem does not run in lane groups'
            self.lib.emu_backward_group(ctypes.byref(ba))
        else:
            (self.lib.emu_backward if flat is None else self.lib.emu_backward_flat)(ctypes.byref(ba))
        return dict(y=fwd['y'], grad=grad_out, lamda=lam_out, status=status, stats=stats,
                    fwd=fwd, tab=tab)
