"""Multi-GPU execution: independent parameter draws sharded across ranks (one process per GPU).

The reference has no in-library parallelism; PyMC runs independent chains in
forked processes.  Here the batch axis is split into contiguous blocks, every rank integrates its
block with no communication, and ONE all-gather per output collects the results
(``torch.distributed``; NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Any, Optional, Tuple


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of rank ``rank``; the first ``n % world`` ranks get one
    extra instance."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError('invalid rank %d / world %d' % (rank, world))
    base, extra = divmod(int(n), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _all_gather_rows(x, counts, group):
    """all-gather of row blocks with (possibly) different row counts: pad to the largest block,
    ``all_gather_into_tensor``, then drop the padding."""
    import torch
    import torch.distributed as dist
    world = len(counts)
    nmax = max(counts)
    pad = x
    if x.shape[0] != nmax:
        pad = torch.zeros((nmax,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[:x.shape[0]] = x
    out = torch.empty((world * nmax,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(c == nmax for c in counts):
        return out
    return torch.cat([out[r * nmax:r * nmax + c] for r, c in enumerate(counts)], dim=0)


def solve_adjoint_sharded(solver: Any, t0: float, tvals, y0, params, grads, *,
                          group: Optional[Any] = None, gather: bool = True):
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
    y, g, lam, status = solver.solve_adjoint_batch(t0, tvals, local(y0), local(params), g_local)
    if not gather or world == 1:
        return y, g, lam, status

    as_numpy = not isinstance(y, torch.Tensor)
    counts = [shard_bounds(B, r, world)[1] - shard_bounds(B, r, world)[0] for r in range(world)]

    def gathered(x):
        t = torch.from_numpy(np.ascontiguousarray(x)) if as_numpy else x
        out = _all_gather_rows(t, counts, group)
        return out.numpy() if as_numpy else out

    # one collective for the trajectories, one for grad | lamda | status packed together
    n_d, n_s = g.shape[1], lam.shape[1]
    if as_numpy:
        small = np.concatenate([g, lam, status[:, None].astype(np.float64)], axis=1)
    else:
        small = torch.cat([g, lam, status[:, None].to(torch.float64)], dim=1)
    y_all = gathered(y)
    small_all = gathered(small)
    g_all, lam_all = small_all[:, :n_d], small_all[:, n_d:n_d + n_s]
    st_all = small_all[:, n_d + n_s]
    st_all = st_all.astype(np.int32) if as_numpy else st_all.to(torch.int32)
    return y_all, g_all, lam_all, st_all
