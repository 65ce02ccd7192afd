"""ctypes binding of the CPU oracle (``oracle/cvodes_port.c``) -- TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module.  The product package ``sunode_b200`` never does.

The oracle integrates with plain-C problem callbacks.  Those come from the C flavour of the
product's generated source (``problem.host_functions``), i.e. the *expressions* are shared with
the product while the *integrator* is independent; hand-derived callbacks for the benchmark
problems live in ``oracle/problems_builtin.c`` and are checked against the generated ones in
``tests/test_codegen.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'liboracle_cvodes.so')
NSTATS = 16

_DP = ctypes.POINTER(ctypes.c_double)
_IP = ctypes.POINTER(ctypes.c_int)
_LP = ctypes.POINTER(ctypes.c_long)


class _Problem(ctypes.Structure):
    _fields_ = [('ns', ctypes.c_int), ('np', ctypes.c_int), ('nd', ctypes.c_int),
                ('rhs', ctypes.c_void_p), ('jac', ctypes.c_void_p), ('adj_rhs', ctypes.c_void_p),
                ('adj_jac', ctypes.c_void_p), ('quad_rhs', ctypes.c_void_p),
                ('sens_rhs', ctypes.c_void_p)]


class _Options(ctypes.Structure):
    _fields_ = [('rtol', ctypes.c_double), ('atol', _DP), ('n_atol', ctypes.c_int),
                ('rtol_b', ctypes.c_double), ('atol_b', ctypes.c_double),
                ('rtol_q', ctypes.c_double), ('atol_q', ctypes.c_double),
                ('mxstep', ctypes.c_int), ('max_retries', ctypes.c_int),
                ('mxstep_b', ctypes.c_int), ('max_retries_b', ctypes.c_int),
                ('hermite', ctypes.c_int), ('constraints', _DP), ('pbar', _DP)]


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe in ``oracle/Makefile``."""
    src = os.path.join(_HERE, 'cvodes_port.c')
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        proc = subprocess.run(['make', '-C', _HERE, '-B' if force else '-s'],
                              capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError('building the oracle failed:\n' + proc.stdout + proc.stderr)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_DP)


class Oracle:
    """CPU reference solver for one problem (any object exposing ``host_functions`` with the
    ``sbh_*`` symbols, or a path to such a shared object together with sizes)."""

    def __init__(self, problem=None, *, host_lib=None, sizes: Optional[Tuple[int, int, int]] = None,
                 rtol=1e-10, atol=1e-10, rtol_b=1e-10, atol_b=1e-10, rtol_q=1e-10, atol_q=1e-10,
                 mxstep=500, max_retries=5, mxstep_b=500, max_retries_b=50, prefix='sbh_',
                 interpolation='polynomial', constraints=None, scaling_factors=None):
        if problem is not None:
            host = problem.host_functions
            self._host = host
            clib = host.lib
            ns, npar, nd = host.gen.n_states, host.gen.n_params, host.gen.n_deriv
        else:
            clib = host_lib if not isinstance(host_lib, str) else ctypes.CDLL(host_lib)
            ns, npar, nd = sizes
        self._clib = clib
        self.ns, self.np, self.nd = ns, npar, nd

        def addr(name):
            return ctypes.cast(getattr(clib, prefix + name), ctypes.c_void_p).value

        self._prob = _Problem(ns, npar, nd, addr('rhs'), addr('jac'), addr('adj_rhs'),
                              addr('adj_jac'), addr('quad_rhs'), addr('sens_rhs'))
        atol_arr = np.atleast_1d(np.asarray(atol, dtype=np.float64)).copy()
        if atol_arr.size not in (1, ns):
            raise ValueError('atol must be a scalar or have one entry per state')
        self._atol = atol_arr
        self._opt = _Options(float(rtol), _dp(atol_arr), int(atol_arr.size), rtol_b, atol_b,
                             rtol_q, atol_q, mxstep, max_retries, mxstep_b, max_retries_b,
                             int(interpolation == 'hermite'), None, None)
        if scaling_factors is not None:
            self._pbar = np.ascontiguousarray(scaling_factors, dtype=np.float64)
            assert self._pbar.shape == (nd,)
            self._opt.pbar = _dp(self._pbar)
        if constraints is not None:
            self._constraints = np.ascontiguousarray(
                np.broadcast_to(np.asarray(constraints, dtype=np.float64), (ns,)))
            self._opt.constraints = _dp(self._constraints)

    # ------------------------------------------------------------------ helpers
    def _prep(self, y0, params, B=None):
        y0 = np.ascontiguousarray(np.atleast_2d(np.asarray(y0, dtype=np.float64)))
        params = np.asarray(params, dtype=np.float64)
        if params.ndim == 1:
            params = params[None, :]
        params = np.ascontiguousarray(params)
        B = max(y0.shape[0], params.shape[0]) if B is None else B
        if y0.shape[0] == 1 and B > 1:
            y0 = np.ascontiguousarray(np.broadcast_to(y0, (B, self.ns)))
        if params.shape[0] == 1 and B > 1:
            params = np.ascontiguousarray(np.broadcast_to(params, (B, self.np)))
        if self.np == 0:
            params = np.zeros((B, 1))
        assert y0.shape == (B, self.ns), y0.shape
        return y0, params, B

    # ------------------------------------------------------------------ solves
    def solve_forward(self, t0, tvals, y0, params, n_threads=0):
        """Batched ``Solver.solve``.  Returns (y_out[B, n_t, ns], status[B], stats[B, 16])."""
        tvals = np.ascontiguousarray(tvals, dtype=np.float64)
        y0, params, B = self._prep(y0, params)
        n_t = len(tvals)
        y_out = np.zeros((B, n_t, self.ns))
        status = np.zeros(B, dtype=np.int32)
        stats = np.zeros((B, NSTATS), dtype=np.int64)
        lib().oracle_solve_forward_batch(
            ctypes.byref(self._prob), ctypes.byref(self._opt), ctypes.c_long(B),
            ctypes.c_double(t0), _dp(tvals), ctypes.c_int(n_t), _dp(y0), _dp(params), _dp(y_out),
            status.ctypes.data_as(_IP), stats.ctypes.data_as(_LP), ctypes.c_int(n_threads))
        return y_out, status, stats

    def solve_forward_sens(self, t0, tvals, y0, params, sens0, n_threads=0):
        """Batched ``Solver(sens_mode=...).solve``.  ``sens0`` is ``[nd, ns]`` (shared) or
        ``[B, nd, ns]``.  Returns (y_out[B, n_t, ns], sens_out[B, n_t, nd, ns], status, stats)."""
        tvals = np.ascontiguousarray(tvals, dtype=np.float64)
        y0, params, B = self._prep(y0, params)
        n_t = len(tvals)
        sens0 = np.ascontiguousarray(sens0, dtype=np.float64)
        shared = int(sens0.ndim == 2)
        assert sens0.shape[-2:] == (self.nd, self.ns)
        y_out = np.zeros((B, n_t, self.ns))
        sens_out = np.zeros((B, n_t, self.nd, self.ns))
        status = np.zeros(B, dtype=np.int32)
        stats = np.zeros((B, NSTATS), dtype=np.int64)
        lib().oracle_solve_forward_sens_batch(
            ctypes.byref(self._prob), ctypes.byref(self._opt), ctypes.c_long(B),
            ctypes.c_double(t0), _dp(tvals), ctypes.c_int(n_t), _dp(y0), _dp(params), _dp(sens0),
            ctypes.c_int(shared), _dp(y_out), _dp(sens_out), status.ctypes.data_as(_IP),
            stats.ctypes.data_as(_LP), ctypes.c_int(n_threads))
        return y_out, sens_out, status, stats

    def solve_adjoint(self, t0, tvals, y0, params, grads, n_threads=0):
        """Batched ``solve_forward`` + ``solve_backward(tvals[-1], t0, tvals, grads, ...)``.

        Returns (y_out, grad_out[B, nd], lamda_out[B, ns], status, stats)."""
        tvals = np.ascontiguousarray(tvals, dtype=np.float64)
        y0, params, B = self._prep(y0, params)
        n_t = len(tvals)
        grads = np.ascontiguousarray(grads, dtype=np.float64)
        shared = int(grads.ndim == 2)
        if shared:
            assert grads.shape == (n_t, self.ns)
        else:
            assert grads.shape == (B, n_t, self.ns)
        y_out = np.zeros((B, n_t, self.ns))
        grad_out = np.zeros((B, max(self.nd, 1)))
        lamda_out = np.zeros((B, self.ns))
        status = np.zeros(B, dtype=np.int32)
        stats = np.zeros((B, NSTATS), dtype=np.int64)
        # grad_out rows have stride nd in C; allocate exactly
        grad_c = np.zeros((B, self.nd)) if self.nd else np.zeros((B, 0))
        gptr = _dp(grad_c) if self.nd else _dp(grad_out)
        lib().oracle_solve_adjoint_batch(
            ctypes.byref(self._prob), ctypes.byref(self._opt), ctypes.c_long(B),
            ctypes.c_double(t0), _dp(tvals), ctypes.c_int(n_t), _dp(y0), _dp(params), _dp(grads),
            ctypes.c_int(shared), _dp(y_out), gptr, _dp(lamda_out),
            status.ctypes.data_as(_IP), stats.ctypes.data_as(_LP), ctypes.c_int(n_threads))
        return y_out, grad_c, lamda_out, status, stats

    def forward_history(self, t0, tvals, y0, params, cap=100000):
        """Stored (t, order, y) points of one adjoint forward pass (for tests)."""
        tvals = np.ascontiguousarray(tvals, dtype=np.float64)
        y0, params, _ = self._prep(y0, params, B=1)
        n_t = len(tvals)
        y_out = np.zeros((n_t, self.ns))
        ht = np.zeros(cap)
        ho = np.zeros(cap, dtype=np.int32)
        hy = np.zeros((cap, self.ns))
        n = ctypes.c_int(0)
        st = lib().oracle_forward_history(
            ctypes.byref(self._prob), ctypes.byref(self._opt), ctypes.c_double(t0), _dp(tvals),
            ctypes.c_int(n_t), _dp(y0), _dp(params), _dp(y_out), ctypes.c_int(cap), _dp(ht),
            ho.ctypes.data_as(_IP), _dp(hy), ctypes.byref(n))
        k = min(n.value, cap)
        return st, y_out, ht[:k], ho[:k], hy[:k]


def max_threads() -> int:
    return int(lib().oracle_max_threads())
