"""``torch.autograd`` binding of the batched adjoint solver.

The reference exposes its gradient through PyTensor only (``sunode/wrappers/as_pytensor.py``);
this is the same contract -- forward solve in ``forward``, ``[-lamda(t0), dL/dparams]`` in
``backward`` (as_pytensor.py:294-308) -- for torch, batched, with all data staying on the device.
"""
from __future__ import annotations

import torch

from ..solver import AdjointSolver


class _SolveIVP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, solver: AdjointSolver, t0: float, tvals, y0, params_deriv, params_fixed):
        subset = solver._problem.params_subset
        B = y0.shape[0]
        full = torch.empty((B, subset.n_items), dtype=torch.float64, device=y0.device)
        if subset.n_subset:
            full[:, torch.as_tensor(subset.subset_flat_index, device=y0.device)] = params_deriv
        if subset.n_items - subset.n_subset:
            idx = torch.as_tensor(subset.remainder_flat_index, device=y0.device)
            full[:, idx] = params_fixed.reshape(1, -1).expand(B, -1) if params_fixed.dim() == 1 else params_fixed
        y0c = y0.detach().contiguous()
        y, status = solver.solve_forward_batch(t0, tvals, y0c, full)
        ctx.solver, ctx.t0, ctx.tvals, ctx.full = solver, t0, tvals, full
        ctx.mark_non_differentiable(status)
        return y, status

    @staticmethod
    def backward(ctx, g_y, _g_status):
        solver = ctx.solver
        # the stored forward pass of the handle is the one made in forward(); a solver shared by
        # several graphs must run forward and backward back to back (as the reference's Ops do)
        grad, lam, _ = solver.solve_backward_batch(ctx.tvals[-1], ctx.t0, ctx.tvals,
                                                   g_y.contiguous(), ctx.full)
        return None, None, None, -lam, grad, None


def solve_ivp(solver: AdjointSolver, t0: float, tvals, y0: torch.Tensor,
              params_deriv: torch.Tensor, params_fixed: torch.Tensor):
    """``y[B, n_t, n_s], status[B] = solve_ivp(...)`` differentiable w.r.t. ``y0[B, n_s]`` and
    ``params_deriv[B, n_deriv]`` (the problem's derivative parameters, in subset order).
    ``params_fixed`` holds the remaining parameters (``[n_fixed]`` shared or ``[B, n_fixed]``)."""
    return _SolveIVP.apply(solver, float(t0), tvals, y0, params_deriv, params_fixed)
