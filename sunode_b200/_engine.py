"""Thin object layer over the C ABI: cubin cache + one solver handle per problem and device.

Everything numeric happens inside ``libsunode_b200.so`` / the JIT-compiled sm_100a kernels; this
module only validates shapes, picks host vs. device pointers and forwards the call.  Arrays may be
numpy arrays (host memory: the library stages them through device buffers it owns) or torch CUDA
tensors (device memory: the kernels work on them in place, asynchronously on torch's current
stream).
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import weakref
from typing import Any, Optional, Tuple

import numpy as np

from . import _build, _lib
from ._cache import atomic_write, cache_dir
from .symode.codegen import GeneratedSource

DEFAULT_ARCH = 'sm_100a'
DEFAULT_BLOCK = int(os.environ.get('SUNODE_B200_BLOCK', '32'))
DEFAULT_MIN_BLOCKS = int(os.environ.get('SUNODE_B200_MIN_BLOCKS', '1'))

_kernel_hash: Optional[str] = None


def kernel_source_hash() -> str:
    """Digest of the device sources + launcher; part of the cubin cache key."""
    global _kernel_hash
    if _kernel_hash is None:
        h = hashlib.sha256()
        # sb_api.cpp: the NVRTC options and the prelude sb_compile adds are part of what a cubin is
        for name in ('sb_args.h', 'sb_bdf.cuh', 'sb_kernels.cuh', 'sb_group.cuh', 'sb_fund.cuh',
                     'sb_api.cpp'):
            with open(os.path.join(_build.CSRC, name), 'rb') as fh:
                h.update(fh.read())
        # the compiler is part of the key: NVRTC versions generate different code (where NVRTC
        # cannot be loaded only cubins of an unknown compiler can be found -- none are shipped)
        try:
            h.update(('nvrtc ' + _lib.nvrtc_version()).encode())
        except _lib.LibraryError:
            h.update(b'nvrtc unavailable')
        _kernel_hash = h.hexdigest()[:16]
    return _kernel_hash


def compile_cubin(gen: GeneratedSource, *, arch: str = DEFAULT_ARCH,
                  block_threads: Optional[int] = None, min_blocks: Optional[int] = None,
                  use_cache: bool = True, defines: Tuple[str, ...] = ()) -> Tuple[bytes, str]:
    """JIT-compile (NVRTC, no GPU needed) the problem + integrator kernels; returns
    ``(cubin, path)``.  Results are cached in-tree keyed by problem, kernel sources and launch
    configuration.  ``defines`` are build options of the kernels that a solver option selects
    (``SB_HERMITE``, ``SB_CONSTRAINTS``); the default build defines none."""
    block = int(block_threads or DEFAULT_BLOCK)
    minb = int(min_blocks or DEFAULT_MIN_BLOCKS)
    # kernel build variants for A/B measurements: SUNODE_B200_DEFINES="SB_INLINE_MATH,SB_FOO=2"
    defines = list(defines) + [d.strip() for d in os.environ.get('SUNODE_B200_DEFINES', '').split(',') if d.strip()]
    prelude = ''.join('#define %s\n' % d.replace('=', ' ', 1) for d in defines)
    tag = ('_' + hashlib.sha256(prelude.encode()).hexdigest()[:8]) if prelude else ''
    path = os.path.join(cache_dir(), 'k_%s_%s_%s_b%d_m%d%s.cubin'
                        % (gen.digest, kernel_source_hash(), arch, block, minb, tag))
    if use_cache and os.path.exists(path):
        with open(path, 'rb') as fh:
            return fh.read(), path
    lib = _lib.lib()
    cubin = ctypes.c_void_p()
    size = ctypes.c_size_t(0)
    log = ctypes.c_void_p()
    code = lib.sb_compile((prelude + gen.cuda).encode(), arch.encode(), block, minb, ctypes.byref(cubin),
                          ctypes.byref(size), ctypes.byref(log))
    log_text = ''
    if log.value:
        log_text = ctypes.string_at(log.value).decode('utf-8', 'replace')
        lib.sb_free(log)
    if code != _lib.SB_OK:
        raise _lib.LibraryError('JIT compilation of the problem kernels failed:\n%s\n%s'
                                % (_lib.last_error(), log_text))
    data = ctypes.string_at(cubin.value, size.value)
    lib.sb_free(cubin)
    atomic_write(path, data)
    return data, path


def _is_torch(x: Any) -> bool:
    return type(x).__module__.split('.')[0] == 'torch'


class _Arg:
    """A validated array argument: pointer + whether it lives on the device."""
    __slots__ = ('ptr', 'device', 'keep')

    def __init__(self, ptr: Optional[int], device: bool, keep: Any):
        self.ptr, self.device, self.keep = ptr, device, keep


def _arg(x: Any, shape: Tuple[int, ...], name: str, dtype=np.float64, writable: bool = False,
         optional: bool = False) -> _Arg:
    if x is None:
        if optional:
            return _Arg(None, False, None)
        raise ValueError('%s is required' % name)
    if _is_torch(x):
        import torch
        want = torch.float64 if dtype == np.float64 else torch.int32
        if x.dtype != want or not x.is_contiguous() or tuple(x.shape) != tuple(shape):
            raise ValueError('%s must be a contiguous %s tensor of shape %s, got %s %s'
                             % (name, want, shape, x.dtype, tuple(x.shape)))
        if not x.is_cuda:
            return _Arg(x.data_ptr(), False, x)
        return _Arg(x.data_ptr(), True, x)
    a = x if writable else np.ascontiguousarray(x, dtype=dtype)
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.c_contiguous:
        raise ValueError('%s must be a C-contiguous %s numpy array' % (name, np.dtype(dtype)))
    if tuple(a.shape) != tuple(shape):
        raise ValueError('%s should have shape %s but has shape %s' % (name, shape, a.shape))
    if writable and not a.flags.writeable:
        raise ValueError('%s must be writable' % name)
    return _Arg(a.ctypes.data if a.size else None, False, a)


def _mem_kind(args, device: Optional[int] = None) -> int:
    kinds = {a.device for a in args if a.ptr is not None}
    if len(kinds) > 1:
        raise ValueError('all arrays of one call must live on the same side (host or device)')
    if kinds == {True} and device is not None:
        # the kernels run in the primary context of the engine's device, on that device's stream
        for a in args:
            if a.ptr is not None and a.keep.device.index != device:
                raise ValueError('tensor on cuda:%s passed to a solver created for cuda:%d'
                                 % (a.keep.device.index, device))
    return _lib.SB_MEM_DEVICE if kinds == {True} else _lib.SB_MEM_HOST


def _stream(mem: int, stream: Optional[int], device: int = 0) -> Optional[int]:
    if stream is not None:
        return stream
    if mem == _lib.SB_MEM_DEVICE:
        import torch
        return torch.cuda.current_stream(device).cuda_stream
    return None


def _destroy(handle_value: int) -> None:
    try:
        _lib.lib().sb_problem_destroy(ctypes.c_void_p(handle_value))
    except Exception:  # noqa: BLE001 - interpreter shutdown
        pass


class Engine:
    """One ``sb_problem`` handle: compiled kernels + device workspace for one problem."""

    def __init__(self, gen: GeneratedSource, *, device: Optional[int] = None,
                 block_threads: Optional[int] = None, min_blocks: Optional[int] = None,
                 arch: str = DEFAULT_ARCH, defines: Tuple[str, ...] = ()):
        self.gen = gen
        self.defines = tuple(defines)
        self.ns, self.np, self.nd = gen.n_states, gen.n_params, gen.n_deriv
        if device is None:
            device = int(os.environ.get('SUNODE_B200_DEVICE', os.environ.get('LOCAL_RANK', '0')))
        self.device = int(device)
        cubin, self.cubin_path = compile_cubin(gen, arch=arch, block_threads=block_threads,
                                               min_blocks=min_blocks, defines=self.defines)
        self._lib = _lib.lib()
        handle = ctypes.c_void_p()
        _lib.check(self._lib.sb_problem_create(ctypes.byref(handle), self.ns, self.np, self.nd,
                                               cubin, len(cubin), self.device))
        self._h = handle
        self._finalizer = weakref.finalize(self, _destroy, handle.value)

    # ------------------------------------------------------------------ configuration
    def set_tolerances(self, rtol: float, atol) -> None:
        atol = np.atleast_1d(np.asarray(atol, dtype=np.float64))
        if atol.ndim != 1 or atol.size not in (1, self.ns):
            raise ValueError('Invalid tolerance.')
        atol = np.ascontiguousarray(atol)
        _lib.check(self._lib.sb_set_tolerances(
            self._h, float(rtol), atol.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            int(atol.size)))

    def set_sens_scaling(self, pbar) -> None:
        if pbar is None:
            _lib.check(self._lib.sb_set_sens_scaling(self._h, None, 0))
            return
        pbar = np.ascontiguousarray(pbar, dtype=np.float64)
        _lib.check(self._lib.sb_set_sens_scaling(
            self._h, pbar.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(pbar.size)))

    def set_tolerances_b(self, rtol: float, atol: float) -> None:
        _lib.check(self._lib.sb_set_tolerances_b(self._h, float(rtol), float(atol)))

    def set_quad_tolerances_b(self, rtol: float, atol: float) -> None:
        _lib.check(self._lib.sb_set_quad_tolerances_b(self._h, float(rtol), float(atol)))

    def set_max_num_steps(self, mxstep: int, max_retries: int) -> None:
        _lib.check(self._lib.sb_set_max_num_steps(self._h, int(mxstep), int(max_retries)))

    def set_max_num_steps_b(self, mxstep: int, max_retries: int) -> None:
        _lib.check(self._lib.sb_set_max_num_steps_b(self._h, int(mxstep), int(max_retries)))

    def set_history_capacity(self, n_steps: int) -> None:
        _lib.check(self._lib.sb_set_history_capacity(self._h, int(n_steps)))

    # ------------------------------------------------------------------ solves
    def forward(self, t0: float, tvals, y0, params, y_out, status, stats=None, *,
                store_history: bool = False, stream: Optional[int] = None) -> None:
        tvals = np.ascontiguousarray(tvals, dtype=np.float64)
        n_t = int(tvals.shape[0])
        B = int(y0.shape[0])
        a_y0 = _arg(y0, (B, self.ns), 'y0')
        a_p = _arg(params, (B, self.np), 'params')
        a_out = _arg(y_out, (B, n_t, self.ns), 'y_out', writable=True)
        a_st = _arg(status, (B,), 'status', dtype=np.int32, writable=True)
        a_stats = _arg(stats, (B, _lib.SB_STATS_PER_INSTANCE), 'stats', dtype=np.int32,
                       writable=True, optional=True)
        mem = _mem_kind([a_y0, a_p, a_out, a_st, a_stats], self.device)
        _lib.check(self._lib.sb_solve_forward(
            self._h, B, float(t0), tvals.ctypes.data, n_t, a_y0.ptr, a_p.ptr, a_out.ptr, a_st.ptr,
            a_stats.ptr, int(bool(store_history)), mem, _stream(mem, stream, self.device)))

    def forward_sens(self, t0: float, tvals, y0, params, sens0, y_out, sens_out, status, stats=None,
                     *, stream: Optional[int] = None) -> None:
        tvals = np.ascontiguousarray(tvals, dtype=np.float64)
        n_t = int(tvals.shape[0])
        B = int(y0.shape[0])
        shared = int(len(sens0.shape) == 2)
        a_y0 = _arg(y0, (B, self.ns), 'y0')
        a_p = _arg(params, (B, self.np), 'params')
        a_s0 = _arg(sens0, (self.nd, self.ns) if shared else (B, self.nd, self.ns), 'sens0')
        a_out = _arg(y_out, (B, n_t, self.ns), 'y_out', writable=True)
        a_so = _arg(sens_out, (B, n_t, self.nd, self.ns), 'sens_out', writable=True)
        a_st = _arg(status, (B,), 'status', dtype=np.int32, writable=True)
        a_stats = _arg(stats, (B, _lib.SB_STATS_PER_INSTANCE), 'stats', dtype=np.int32,
                       writable=True, optional=True)
        mem = _mem_kind([a_y0, a_p, a_s0, a_out, a_so, a_st, a_stats], self.device)
        _lib.check(self._lib.sb_solve_forward_sens(
            self._h, B, float(t0), tvals.ctypes.data, n_t, a_y0.ptr, a_p.ptr, a_s0.ptr, shared,
            a_out.ptr, a_so.ptr, a_st.ptr, a_stats.ptr, mem, _stream(mem, stream, self.device)))

    def backward(self, t_start: float, t_end: float, tvals, params, grads, grad_out, lamda_out,
                 status, stats=None, *, lamda_all=None, quad_all=None,
                 stream: Optional[int] = None) -> None:
        tvals = np.ascontiguousarray(tvals, dtype=np.float64)
        n_t = int(tvals.shape[0])
        B = int(lamda_out.shape[0])
        a_la = _arg(lamda_all, (B, n_t, self.ns), 'lamda_all', writable=True, optional=True)
        a_qa = _arg(quad_all, (B, n_t, self.nd), 'quad_all', writable=True, optional=True)
        shared = int(len(grads.shape) == 2)
        a_p = _arg(params, (B, self.np), 'params', optional=True)
        a_g = _arg(grads, (n_t, self.ns) if shared else (B, n_t, self.ns), 'grads')
        a_go = _arg(grad_out, (B, self.nd), 'grad_out', writable=True)
        a_lo = _arg(lamda_out, (B, self.ns), 'lamda_out', writable=True)
        a_st = _arg(status, (B,), 'status', dtype=np.int32, writable=True)
        a_stats = _arg(stats, (B, _lib.SB_STATS_PER_INSTANCE), 'stats', dtype=np.int32,
                       writable=True, optional=True)
        mem = _mem_kind([a_p, a_g, a_go, a_lo, a_st, a_stats, a_la, a_qa], self.device)
        if a_la.ptr is not None or a_qa.ptr is not None:
            _lib.check(self._lib.sb_set_backward_trace(self._h, a_la.ptr, a_qa.ptr))
        _lib.check(self._lib.sb_solve_backward(
            self._h, B, float(t_start), float(t_end), tvals.ctypes.data, n_t, a_p.ptr, a_g.ptr,
            shared, a_go.ptr, a_lo.ptr, a_st.ptr, a_stats.ptr, mem, _stream(mem, stream, self.device)))

    def adjoint(self, t0: float, tvals, y0, params, grads, y_out, grad_out, lamda_out, status,
                stats_fwd=None, stats_bwd=None, *, stream: Optional[int] = None) -> None:
        tvals = np.ascontiguousarray(tvals, dtype=np.float64)
        n_t = int(tvals.shape[0])
        B = int(y0.shape[0])
        shared = int(len(grads.shape) == 2)
        a_y0 = _arg(y0, (B, self.ns), 'y0')
        a_p = _arg(params, (B, self.np), 'params')
        a_g = _arg(grads, (n_t, self.ns) if shared else (B, n_t, self.ns), 'grads')
        a_out = _arg(y_out, (B, n_t, self.ns), 'y_out', writable=True)
        a_go = _arg(grad_out, (B, self.nd), 'grad_out', writable=True)
        a_lo = _arg(lamda_out, (B, self.ns), 'lamda_out', writable=True)
        a_st = _arg(status, (B,), 'status', dtype=np.int32, writable=True)
        a_sf = _arg(stats_fwd, (B, _lib.SB_STATS_PER_INSTANCE), 'stats_fwd', dtype=np.int32,
                    writable=True, optional=True)
        a_sb = _arg(stats_bwd, (B, _lib.SB_STATS_PER_INSTANCE), 'stats_bwd', dtype=np.int32,
                    writable=True, optional=True)
        mem = _mem_kind([a_y0, a_p, a_g, a_out, a_go, a_lo, a_st, a_sf, a_sb], self.device)
        _lib.check(self._lib.sb_solve_adjoint(
            self._h, B, float(t0), tvals.ctypes.data, n_t, a_y0.ptr, a_p.ptr, a_g.ptr, shared,
            a_out.ptr, a_go.ptr, a_lo.ptr, a_st.ptr, a_sf.ptr, a_sb.ptr, mem,
            _stream(mem, stream, self.device)))

    def eval(self, kind: int, t, y, params, lam, out, *, params_shared: bool = False,
             stream: Optional[int] = None) -> None:
        n = int(y.shape[0])
        n_out = {0: self.ns, 1: self.ns * self.ns, 2: self.ns, 3: self.nd, 4: self.ns * self.ns}[kind]
        a_t = _arg(t, (n,), 't')
        a_y = _arg(y, (n, self.ns), 'y')
        a_p = _arg(params, (self.np,) if params_shared else (n, self.np), 'params')
        a_l = _arg(lam, (n, self.ns), 'lam', optional=kind in (0, 1, 4))
        a_o = _arg(out, (n, n_out), 'out', writable=True)
        mem = _mem_kind([a_t, a_y, a_p, a_l, a_o], self.device)
        _lib.check(self._lib.sb_eval(self._h, int(kind), n, a_t.ptr, a_y.ptr, a_p.ptr,
                                     int(bool(params_shared)), a_l.ptr, a_o.ptr, mem,
                                     _stream(mem, stream, self.device)))

    def set_workspace_limit(self, n_bytes: int) -> None:
        _lib.check(self._lib.sb_set_workspace_limit(self._h, int(n_bytes)))

    def last_chunks(self) -> int:
        return int(self._lib.sb_last_chunks(self._h))

    def forward_fail_index(self, B: int) -> np.ndarray:
        out = np.empty((int(B),), dtype=np.int32)
        _lib.check(self._lib.sb_forward_fail_index(self._h, int(B), out.ctypes.data))
        return out

    # ------------------------------------------------------------------ introspection
    def synchronize(self) -> None:
        _lib.check(self._lib.sb_synchronize(self._h))

    def last_kernel_ms(self) -> Tuple[float, float, float]:
        f, t, b = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
        _lib.check(self._lib.sb_last_kernel_ms(self._h, ctypes.byref(f), ctypes.byref(t),
                                               ctypes.byref(b)))
        return f.value, t.value, b.value

    def launch_count(self) -> int:
        return int(self._lib.sb_launch_count(self._h))

    def kernel_info(self) -> dict:
        vals = [ctypes.c_int() for _ in range(6)]
        _lib.check(self._lib.sb_kernel_info(self._h, *[ctypes.byref(v) for v in vals]))
        keys = ('regs_fwd', 'regs_bwd', 'blocks_per_sm_fwd', 'blocks_per_sm_bwd', 'block_threads',
                'sm_count')
        return {k: v.value for k, v in zip(keys, vals)}


def lanes_per_instance(n_states: int) -> int:
    """Lanes of a warp that integrate one instance in the backward kernels: ``SB_GROUP_SIZE`` of
    ``csrc/sb_args.h`` (one lane up to 4 states, then a power-of-two group with two state
    components per lane, see ``csrc/sb_group.cuh``)."""
    if 'SB_NO_GROUP' in os.environ.get('SUNODE_B200_DEFINES', ''):
        return 1
    if n_states < 5 or n_states > 64:
        return 1
    return 4 if n_states <= 8 else 8 if n_states <= 16 else 16 if n_states <= 32 else 32


FLAT_FWD_STEPS_PER_TVAL = 8     # SB_FLAT_FWD_STEPS_PER_TVAL of csrc/sb_api.cpp


class PinnedBuffer:
    """Page-locked host array (``sb_host_alloc``) for asynchronous staging."""

    def __init__(self, shape, dtype=np.float64):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape, dtype=np.int64)) * dtype.itemsize
        ptr = ctypes.c_void_p()
        _lib.check(_lib.lib().sb_host_alloc(ctypes.byref(ptr), nbytes))
        self._ptr = ptr
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=nbytes // dtype.itemsize).reshape(self.shape)
        self._finalizer = weakref.finalize(self, PinnedBuffer._free, ptr.value)

    @staticmethod
    def _free(ptr_value: int) -> None:
        try:
            _lib.lib().sb_host_free(ctypes.c_void_p(ptr_value))
        except Exception:  # noqa: BLE001
            pass
