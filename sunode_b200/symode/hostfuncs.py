"""Host-side evaluators of the generated problem functions.

The C flavour of the generated source is compiled with the system C compiler into a small shared
object and called through ctypes.  It backs ``SympyProblem.make_rhs()`` & co. (which the reference
implements with numba, ``sunode/symode/problem.py:251-465``) and provides plain C function
pointers with the calling convention documented in :mod:`.codegen`.  It is *not* used by the GPU
solve path: the CUDA flavour of the same source is compiled into the kernels.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

from .._cache import atomic_write, cache_dir
from .codegen import GeneratedSource

_DP = ctypes.POINTER(ctypes.c_double)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_DP)


def compile_host_module(gen: GeneratedSource, cc: Optional[str] = None) -> str:
    """Compile (or find in the cache) the host shared object; returns its path."""
    out = os.path.join(cache_dir(), 'host_%s.so' % gen.digest)
    if os.path.exists(out):
        return out
    src = os.path.join(cache_dir(), 'host_%s.c' % gen.digest)
    atomic_write(src, gen.c.encode())
    tmp = '%s.tmp%d' % (out, os.getpid())
    cmd = [cc or os.environ.get('CC', 'gcc'), '-O2', '-fPIC', '-shared', '-std=c99',
           '-fno-fast-math', '-o', tmp, src, '-lm']
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('Compiling generated C failed:\n%s\n%s' % (' '.join(cmd), proc.stderr))
    os.replace(tmp, out)
    return out


class HostFunctions:
    """ctypes view of one compiled problem module."""

    def __init__(self, gen: GeneratedSource):
        self.gen = gen
        self.path = compile_host_module(gen)
        self.lib = ctypes.CDLL(self.path)
        d = ctypes.c_double
        self.lib.sbh_rhs.argtypes = [d, _DP, _DP, _DP]
        self.lib.sbh_jac.argtypes = [d, _DP, _DP, _DP]
        self.lib.sbh_adj_jac.argtypes = [d, _DP, _DP, _DP]
        self.lib.sbh_adj_rhs.argtypes = [d, _DP, _DP, _DP, _DP]
        self.lib.sbh_quad_rhs.argtypes = [d, _DP, _DP, _DP, _DP]
        self.lib.sbh_sens_rhs.argtypes = [d, _DP, _DP, _DP, _DP]
        for name in ('sbh_rhs', 'sbh_jac', 'sbh_adj_jac', 'sbh_adj_rhs', 'sbh_quad_rhs',
                     'sbh_sens_rhs'):
            getattr(self.lib, name).restype = ctypes.c_int

    @staticmethod
    def _in(a, n) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        if a.size != n:
            raise ValueError('Expected %d values, got %d' % (n, a.size))
        if a.size == 0:
            a = np.zeros(1)
        return a

    def _call3(self, fn, t, y, p, out, n_out):
        g = self.gen
        y = self._in(y, g.n_states)
        p = self._in(p, g.n_params)
        buf = np.zeros(max(n_out, 1))
        flag = fn(t, _ptr(y), _ptr(p), _ptr(buf))
        np.asarray(out).reshape(-1)[...] = buf[:n_out]
        return int(flag)

    def _call4(self, fn, t, y, v, nv, p, out, n_out):
        g = self.gen
        y = self._in(y, g.n_states)
        v = self._in(v, nv)
        p = self._in(p, g.n_params)
        buf = np.zeros(max(n_out, 1))
        flag = fn(t, _ptr(y), _ptr(v), _ptr(p), _ptr(buf))
        np.asarray(out).reshape(-1)[...] = buf[:n_out]
        return int(flag)

    def rhs(self, t, y, p, out):
        return self._call3(self.lib.sbh_rhs, t, y, p, out, self.gen.n_states)

    def jac(self, t, y, p, out):
        """Column-major ``out[i + NS*j]``."""
        return self._call3(self.lib.sbh_jac, t, y, p, out, self.gen.n_states ** 2)

    def adj_jac(self, t, y, p, out):
        return self._call3(self.lib.sbh_adj_jac, t, y, p, out, self.gen.n_states ** 2)

    def adj_rhs(self, t, y, lam, p, out):
        g = self.gen
        return self._call4(self.lib.sbh_adj_rhs, t, y, lam, g.n_states, p, out, g.n_states)

    def quad_rhs(self, t, y, lam, p, out):
        g = self.gen
        return self._call4(self.lib.sbh_quad_rhs, t, y, lam, g.n_states, p, out, g.n_deriv)

    def sens_rhs(self, t, y, s, p, out):
        g = self.gen
        return self._call4(self.lib.sbh_sens_rhs, t, y, s, g.n_deriv * g.n_states, p, out,
                           g.n_deriv * g.n_states)
