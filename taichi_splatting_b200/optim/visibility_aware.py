"""Visibility-aware Adam / LaProp: fractional optimisers whose per-point step weight is the ratio of a point's
current visibility to its running visibility (behaviour of taichi_splatting/optim/visibility_aware.py:11-125).

One step over the visible points (`indexes`, their `visibility` from the rasteriser):
  1. gs_optim_update_visibility_f32 -- running_vis <- power-4 mean of (visibility, running_vis) with weight vis_beta;
     weight = visibility / running_vis; total_weight += weight        (one launch; upstream: ~10 torch passes)
  2. per parameter group gs_optim_step_f32 with gradients divided by (visibility + vis_smooth) on the fly
     (upstream materialises a scaled (N, D) copy of every gradient first).
"""
from typing import Optional

import torch

from .. import _lib
from . import fractional as fr


def get_running_vis(state: dict, n: int, device: torch.device) -> torch.Tensor:
  return state.setdefault('running_vis', torch.zeros((n,), device=device, dtype=torch.float32))


def update_visibility(running_vis: torch.Tensor, visibility: torch.Tensor, indexes: torch.Tensor,
                      total_weight: torch.Tensor, beta: float = 0.9, eps: float = 1e-12) -> torch.Tensor:
  """Advances running_vis[indexes] and total_weight[indexes] in place; returns the step weights (M,).
  (Upstream's function of this name leaves the total_weight update to its caller, :90-91; here both are one kernel.)"""
  _lib.require_cuda(running_vis=running_vis, visibility=visibility, indexes=indexes, total_weight=total_weight)
  vis = visibility.to(torch.float32).contiguous()
  weight = torch.empty_like(vis)
  indexes_c = indexes.contiguous()   # named: the pointer must not outlive a temporary copy
  _lib.call("gs_optim_update_visibility_f32", _lib.ptr(running_vis), _lib.ptr(vis), _lib.ptr(indexes_c),
            _lib.ptr(total_weight), float(beta), float(eps), indexes.shape[0], _lib.ptr(weight),
            _lib.stream_ptr(running_vis.device))
  return weight


class VisibilityOptimizer(torch.optim.Optimizer):
  """`kernels`: fractional.ADAM | fractional.LAPROP (or a reference-style kernel module)."""

  def __init__(self, kernels, params: list, lr=0.001, betas=(0.9, 0.999), eps=1e-16, vis_beta=0.9,
               vis_smooth: float = 0.01, bias_correction=True, grad_clip: Optional[float] = None):
    assert 0.0 <= vis_beta < 1.0, f"Invalid visibility beta: {vis_beta}"
    self.kernels = fr.resolve_algorithm(kernels)
    self.vis_beta, self.vis_smooth = vis_beta, vis_smooth
    super().__init__(params, fr.group_defaults(lr, betas, eps, bias_correction, grad_clip))

  @torch.no_grad()
  def step(self, indexes: torch.Tensor, visibility: torch.Tensor, basis: Optional[torch.Tensor] = None):
    assert visibility.shape == indexes.shape, f"shape mismatch {visibility.shape} != {indexes.shape}"
    groups = [fr.make_group(g, self.state) for g in self.param_groups]
    shared, n, device = groups[0].state, groups[0].num_points, visibility.device
    total_weight = fr.get_total_weight(shared, n, device=device)
    vis = visibility.to(torch.float32).contiguous()
    weight = update_visibility(get_running_vis(shared, n, device), vis, indexes, total_weight, self.vis_beta)
    for group in groups:
      if group.grad is not None:
        assert group.num_points == n, f"param shape {group.num_points} != {n}"
        fr.apply_step(group, weight, indexes, total_weight, self.kernels, basis, grad_scale=vis,
                      grad_smooth=self.vis_smooth)


def _visibility_optimiser(name: str, algorithm: int):
  def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-16, vis_beta=0.5, vis_smooth: float = 0.01,
               bias_correction=True, grad_clip: Optional[float] = None):
    VisibilityOptimizer.__init__(self, algorithm, params, lr=lr, betas=betas, eps=eps, vis_beta=vis_beta,
                                 vis_smooth=vis_smooth, bias_correction=bias_correction, grad_clip=grad_clip)
  return type(name, (VisibilityOptimizer,), {"__init__": __init__, "__module__": __name__,
                                             "__doc__": f"VisibilityOptimizer with the {('Adam', 'LaProp')[algorithm]} step."})


VisibilityAwareAdam = _visibility_optimiser("VisibilityAwareAdam", fr.ADAM)
VisibilityAwareLaProp = _visibility_optimiser("VisibilityAwareLaProp", fr.LAPROP)
