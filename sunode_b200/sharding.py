"""Multi-GPU execution: independent parameter draws sharded across ranks (one process per GPU).

The reference has no in-library parallelism; PyMC runs independent chains in
forked processes.  Here the batch axis is split into contiguous blocks, every rank integrates its
block with no communication, and ONE all-gather per output collects the results
(``torch.distributed``; NCCL over NVLink on GPUs, gloo in the CPU tests).

The trajectories are final as soon as the forward kernel is done -- one millisecond into a
seventeen-millisecond step -- so their all-gather (the large one: ``B * n_t * n_s`` doubles per
rank) is issued right there, asynchronously, and travels over NVLink underneath the backward
kernels; only the small ``grad | lamda | status`` gather follows the backward pass.

Transport of the large gather on GPUs: the forward kernel writes its trajectories into a buffer
in *symmetric memory* (``torch.distributed._symmetric_memory``: the same allocation on every
rank, mapped into every peer's address space over NVLink); after a device-side barrier every
rank PULLS the blocks of its peers with plain peer-to-peer copies, which the copy engines carry
out -- no SM is taken from the persistent backward kernel, which an NCCL all-gather kernel does
(measured on 8 B200s: backward kernel 15.46 -> 15.71 ms under the NCCL ring).  Where symmetric
memory is not available (no peer access, CPU tensors) the gather is an NCCL / gloo
``all_gather_into_tensor``.
"""
from __future__ import annotations

import os
from typing import Any, List, Optional, Tuple


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of rank ``rank``; the first ``n % world`` ranks get one
    extra instance."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError('invalid rank %d / world %d' % (rank, world))
    base, extra = divmod(int(n), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _Gather:
    """An all-gather of row blocks with (possibly) different row counts, in flight: blocks are
    padded to the largest one, ``all_gather_into_tensor(async_op=True)``; :meth:`result` makes the
    current stream (CUDA) or the caller (CPU) wait and drops the padding."""

    def __init__(self, x, counts: List[int], group, out=None):
        import torch
        import torch.distributed as dist
        self.counts, self.nmax = counts, max(counts)
        world = len(counts)
        pad = x
        if x.shape[0] != self.nmax:
            pad = torch.zeros((self.nmax,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            pad[:x.shape[0]] = x
        self._keep = pad.contiguous()
        shape = (world * self.nmax,) + tuple(x.shape[1:])
        if out is None or tuple(out.shape) != shape:
            out = torch.empty(shape, dtype=x.dtype, device=x.device)
        self.out = out
        self.work = dist.all_gather_into_tensor(self.out, self._keep, group=group, async_op=True)

    def result(self):
        import torch
        self.work.wait()
        if all(c == self.nmax for c in self.counts):
            return self.out
        return torch.cat([self.out[r * self.nmax:r * self.nmax + c]
                          for r, c in enumerate(self.counts)], dim=0)


class _PeerGather:
    """The all-gather of equal row blocks held in symmetric memory, in flight on a side stream:
    a device-side barrier (every rank's block is final), then one peer-to-peer copy per rank into
    the local output -- the copy engines pull over NVLink, no SM is involved.  ``ready`` is an
    event recorded on the producing stream right after the forward kernel; the gather is enqueued
    AFTER the backward kernels so that its host-side cost (a dozen enqueues at 8 ranks) does not
    delay their launch.  :meth:`result` makes the current stream wait for the copies.

    The symmetric block may be overwritten again once every peer has pulled it.  No second barrier
    is spent on that: the caller issues its trailing (NCCL) collective only after :meth:`result`,
    so a rank's contribution to that collective implies its pulls are done, and the collective's
    completion on the owner implies every peer's contribution."""

    def __init__(self, x, handle, world: int, ready, out=None):
        import torch
        rows = x.shape[0]
        shape = (world * rows,) + tuple(x.shape[1:])
        if out is None or tuple(out.shape) != shape:
            out = torch.empty(shape, dtype=x.dtype, device=x.device)
        self.out = out
        self.main = torch.cuda.current_stream(x.device)
        self.side = _copy_stream(x.device, 'gather')
        self.side.wait_event(ready)
        rank = handle.rank
        with torch.cuda.stream(self.side):
            # (bounded: a rank that died must not leave its peers spinning on the device)
            handle.barrier(timeout_ms=120000)
            for step in range(world):
                r = (rank - step) % world                       # own block first, then round the ring
                src = x if r == rank else handle.get_buffer(r, tuple(x.shape), x.dtype)
                out[r * rows:(r + 1) * rows].copy_(src, non_blocking=True)

    def result(self):
        self.main.wait_stream(self.side)
        return self.out


_symm: dict = {}
_last_transport: Optional[str] = None


def last_transport() -> Optional[str]:
    """``'p2p'`` or ``'nccl'``: how the last overlapped trajectory gather of this process travelled."""
    return _last_transport


def symmetric_rows(shape, device, group=None):
    """``(tensor, handle)``: a float64 tensor of ``shape`` in symmetric memory, the same on every
    rank of ``group`` (collective call; cached per shape), or ``(None, None)`` where symmetric
    memory cannot be set up -- the callers then use the NCCL transport."""
    import torch
    import torch.distributed as dist
    key = (tuple(shape), torch.device(device).index, id(group))
    if key not in _symm:
        try:
            import torch.distributed._symmetric_memory as symm_mem
            pg = group if group is not None else dist.group.WORLD
            t = symm_mem.empty(tuple(shape), dtype=torch.float64, device=torch.device(device))
            _symm[key] = (t, symm_mem.rendezvous(t, pg))
        except Exception:  # noqa: BLE001 - no peer access / unsupported build: NCCL transport
            _symm[key] = (None, None)
        # all ranks must agree, or the collectives of the two transports would be mismatched
        ok = torch.tensor([int(_symm[key][0] is not None)], device=torch.device(device))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            _symm[key] = (None, None)
    return _symm[key]


_copy_streams: dict = {}


def _copy_stream(device, purpose='download'):
    import torch
    key = (torch.device(device).index, purpose)
    if key not in _copy_streams:
        # high priority: the gather's barrier kernel (one block) must get its slot while the table
        # kernel's short blocks turn over -- once the persistent backward kernel has filled every
        # SM nothing else is scheduled before its tail (seen: copies done at 16.6 ms instead of 1.4)
        _copy_streams[key] = torch.cuda.Stream(device=device, priority=-1)
    return _copy_streams[key]


def solve_adjoint_gathered(solver: Any, t0: float, tvals, y0, params, grads, counts: List[int], *,
                           group: Optional[Any] = None, y_out=None, grad_out=None, lamda_out=None,
                           status=None, y_all=None, small_all=None, overlap: Optional[bool] = None,
                           host_out: Optional[dict] = None, transport: Optional[str] = None):
    """This rank's shard (``y0[B_r, n_s]`` ...) solved forward + adjoint, results of ALL ranks
    returned: ``(y_all, grad_all, lamda_all, status_all)``; ``counts[r]`` = instances of rank r.

    ``overlap`` (default: whenever the solver offers the split calls and the shard lives on a
    GPU): forward pass, then the trajectories' all-gather asynchronously, then the backward pass,
    so that the collective runs underneath the backward kernels.  Otherwise one fused
    ``solve_adjoint_batch`` followed by both gathers.  Results are identical either way.

    Transport of the trajectories' gather on GPUs: peer-to-peer pulls from symmetric memory by
    the copy engines when ``transport`` is ``'p2p'`` or (default) ``None`` and symmetric memory can
    be set up, else NCCL; ``y_out``, if given, must then come from :func:`symmetric_rows` (the
    forward kernel writes into it).  ``last_transport()`` tells which one ran.

    ``host_out`` (overlapped path only): pinned CPU tensors ``{'y', 'g', 'l', 'st'}`` that receive
    this rank's OWN results; the trajectories' download runs on a copy stream underneath the
    backward kernels too.  The caller synchronises the current stream before reading them."""
    import numpy as np
    import torch

    as_numpy = not isinstance(y0, torch.Tensor)
    split = hasattr(solver, 'solve_forward_batch') and hasattr(solver, 'solve_backward_batch')
    if overlap is None:
        overlap = split and not as_numpy and y0.is_cuda

    def tensor(x):
        return torch.from_numpy(np.ascontiguousarray(x)) if as_numpy else x

    global _last_transport
    if transport is None:
        transport = os.environ.get('SUNODE_B200_GATHER') or None       # 'nccl' | 'p2p' (A/B runs)
    if overlap:
        # the forward kernel writes the trajectories straight into symmetric memory when it can
        handle = None
        if (transport in (None, 'p2p') and not as_numpy and y0.is_cuda and len(set(counts)) == 1
                and hasattr(solver, '_problem')):
            shape = (int(y0.shape[0]), len(tvals), solver._problem.n_states)
            if y_out is not None:
                # a caller-supplied output buffer is used as it is: peer pulls only if it IS a
                # symmetric buffer (symmetric_rows), else the NCCL transport
                handle = next((h for t, h in _symm.values()
                               if t is not None and t.data_ptr() == y_out.data_ptr()
                               and tuple(t.shape) == tuple(y_out.shape)), None)
            else:
                y_out, handle = symmetric_rows(shape, y0.device, group)
        if transport == 'p2p' and handle is None:
            raise RuntimeError('symmetric memory is not available for the p2p transport')
        y, st_f = solver.solve_forward_batch(t0, tvals, y0, params, y_out=y_out)
        pending_y = None
        if handle is not None:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(y.device))            # y is final here
            _last_transport = 'p2p'
        else:
            # an NCCL kernel needs SMs: it is launched before the persistent backward kernel
            # fills them, and runs next to it
            pending_y = _Gather(tensor(y), counts, group, out=y_all)
            _last_transport = 'nccl'
        if host_out is not None:
            main = torch.cuda.current_stream(y.device)
            side = _copy_stream(y.device)
            side.wait_stream(main)                                       # y is final after the forward kernel
            with torch.cuda.stream(side):
                host_out['y'].copy_(y, non_blocking=True)
        g, lam, status = solver.solve_backward_batch(tvals[-1], t0, tvals, grads, grad_out=grad_out,
                                                     lamda_out=lamda_out, status=status)
        if pending_y is None:
            pending_y = _PeerGather(y, handle, len(counts), ready, out=y_all)   # copy engines, under the backward pass
        if host_out is not None:
            host_out['g'].copy_(g, non_blocking=True)
            host_out['l'].copy_(lam, non_blocking=True)
            host_out['st'].copy_(status, non_blocking=True)
            main.wait_stream(side)
    else:
        kw = {}
        if y_out is not None:
            kw = dict(y_out=y_out, grad_out=grad_out, lamda_out=lamda_out, status=status)
        y, g, lam, status = solver.solve_adjoint_batch(t0, tvals, y0, params, grads, **kw)
        pending_y = _Gather(tensor(y), counts, group, out=y_all)
    # grad | lamda | status packed into one small collective
    n_d, n_s = g.shape[1], lam.shape[1]
    if as_numpy:
        small = torch.from_numpy(np.concatenate([g, lam, status[:, None].astype(np.float64)], axis=1))
    else:
        small = torch.cat([g, lam, status[:, None].to(torch.float64)], dim=1)
    y_all = pending_y.result()          # (before the small collective: see _PeerGather)
    small_all = _Gather(small, counts, group, out=small_all).result()
    g_all, lam_all = small_all[:, :n_d], small_all[:, n_d:n_d + n_s]
    st_all = small_all[:, n_d + n_s].to(torch.int32)
    if as_numpy:
        return y_all.numpy(), g_all.numpy(), lam_all.numpy(), st_all.numpy()
    return y_all, g_all, lam_all, st_all


def solve_adjoint_sharded(solver: Any, t0: float, tvals, y0, params, grads, *,
                          group: Optional[Any] = None, gather: bool = True,
                          overlap: Optional[bool] = None):
    """Forward + adjoint solve of a GLOBAL batch, sharded over the ranks of ``group``.

    Every rank passes the same global ``y0[B, n_s]`` / ``params[B, n_all]`` (torch tensors on its
    own device, or numpy arrays) and, for per-instance cotangents, ``grads[B, n_t, n_s]``; a 2-D
    ``grads`` is shared by all instances.  Returns ``(y_out, grad_out, lamda_out, status)`` for
    the global batch on every rank (``gather=True``) or for the local shard only."""
    import numpy as np
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = int(y0.shape[0])
    lo, hi = shard_bounds(B, rank, world)

    def local(x):
        part = x[lo:hi]
        return part.contiguous() if isinstance(part, torch.Tensor) else np.ascontiguousarray(part)

    g_local = grads if len(grads.shape) == 2 else local(grads)
    if not gather or world == 1:
        return solver.solve_adjoint_batch(t0, tvals, local(y0), local(params), g_local)
    counts = [shard_bounds(B, r, world)[1] - shard_bounds(B, r, world)[0] for r in range(world)]
    return solve_adjoint_gathered(solver, t0, tvals, local(y0), local(params), g_local, counts,
                                  group=group, overlap=overlap)
