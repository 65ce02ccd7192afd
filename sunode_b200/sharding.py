"""Multi-GPU execution: independent parameter draws sharded across ranks (one process per GPU).

The reference has no in-library parallelism; PyMC runs independent chains in
forked processes.  Here the batch axis is split into contiguous blocks, every rank integrates its
block with no communication, and ONE all-gather per output collects the results
(``torch.distributed``; NCCL over NVLink on GPUs, gloo in the CPU tests).

The trajectories are final as soon as the forward kernel is done -- one millisecond into a
twenty-millisecond step -- so their all-gather (the large one: ``B * n_t * n_s`` doubles per rank)
is issued right there, asynchronously, and travels over NVLink underneath the backward kernels;
only the small ``grad | lamda | status`` gather follows the backward pass.
"""
from __future__ import annotations

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


_copy_streams: dict = {}


def _copy_stream(device):
    import torch
    key = torch.device(device).index
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=device)
    return _copy_streams[key]


def solve_adjoint_gathered(solver: Any, t0: float, tvals, y0, params, grads, counts: List[int], *,
                           group: Optional[Any] = None, y_out=None, grad_out=None, lamda_out=None,
                           status=None, y_all=None, small_all=None, overlap: Optional[bool] = None,
                           host_out: Optional[dict] = None):
    """This rank's shard (``y0[B_r, n_s]`` ...) solved forward + adjoint, results of ALL ranks
    returned: ``(y_all, grad_all, lamda_all, status_all)``; ``counts[r]`` = instances of rank r.

    ``overlap`` (default: whenever the solver offers the split calls and the shard lives on a
    GPU): forward pass, then the trajectories' all-gather asynchronously, then the backward pass,
    so that the collective runs underneath the backward kernels.  Otherwise one fused
    ``solve_adjoint_batch`` followed by both gathers.  Results are identical either way.

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

    if overlap:
        y, st_f = solver.solve_forward_batch(t0, tvals, y0, params, y_out=y_out)
        pending_y = _Gather(tensor(y), counts, group, out=y_all)         # travels under the backward pass
        if host_out is not None:
            main = torch.cuda.current_stream(y.device)
            side = _copy_stream(y.device)
            side.wait_stream(main)                                       # y is final after the forward kernel
            with torch.cuda.stream(side):
                host_out['y'].copy_(y, non_blocking=True)
        g, lam, status = solver.solve_backward_batch(tvals[-1], t0, tvals, grads, grad_out=grad_out,
                                                     lamda_out=lamda_out, status=status)
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
    small_all = _Gather(small, counts, group, out=small_all).result()
    y_all = pending_y.result()
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
