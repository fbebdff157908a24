"""Visibility-aware Adam / LaProp (reference: taichi_splatting/optim/visibility_aware.py:11-125).

step(indexes, visibility, basis=None): the running visibility of the visible points is advanced
(gs_optim_update_visibility_f32), its ratio to the current visibility is the fractional step weight, gradients are
divided by (visibility + vis_smooth), then the fractional step runs (one fused launch per parameter group)."""
from typing import Optional

import torch

from .. import _lib
from .fractional import ADAM, LAPROP, apply_step, get_total_weight, make_group


def get_running_vis(state: dict, n: int, device: torch.device):
  if 'running_vis' not in state:
    state['running_vis'] = torch.zeros((n,), device=device, dtype=torch.float32)
  return state['running_vis']


def update_visibility(running_vis: torch.Tensor, visibility: torch.Tensor, indexes: torch.Tensor,
                      total_weight: torch.Tensor, beta: float = 0.9, eps: float = 1e-12):
  """Reference :37-48 plus the step-count update of :90-91 (total_weight[indexes] += weight) -> weight (M,)."""
  _lib.require_cuda(running_vis=running_vis, visibility=visibility, indexes=indexes, total_weight=total_weight)
  weight = torch.empty_like(visibility, dtype=torch.float32)
  ptr = _lib.ptr
  _lib.call("gs_optim_update_visibility_f32", ptr(running_vis), ptr(visibility.to(torch.float32).contiguous()),
            ptr(indexes.contiguous()), ptr(total_weight), float(beta), float(eps), indexes.shape[0], ptr(weight),
            _lib.stream_ptr(running_vis.device))
  return weight


class VisibilityOptimizer(torch.optim.Optimizer):
  def __init__(self, kernels, params: list, lr=0.001, betas=(0.9, 0.999), eps=1e-16, vis_beta=0.9,
               vis_smooth: float = 0.01, bias_correction=True, grad_clip: Optional[float] = None):
    assert lr > 0, f"Invalid learning rate: {lr}"
    assert eps > 0, f"Invalid epsilon: {eps}"
    assert 0.0 <= betas[0] < 1.0, f"Invalid beta1: {betas[0]}"
    assert 0.0 <= betas[1] < 1.0, f"Invalid beta2: {betas[1]}"
    assert 0.0 <= vis_beta < 1.0, f"Invalid visibility beta: {vis_beta}"
    assert kernels in (ADAM, LAPROP)
    defaults = dict(lr=lr, betas=betas, eps=eps, mask_lr=None, point_lr=None, type="scalar",
                    bias_correction=bias_correction, clip=grad_clip)
    self.vis_beta = vis_beta
    self.vis_smooth = vis_smooth
    self.kernels = kernels
    super().__init__(params, defaults)

  @torch.no_grad()
  def step(self, indexes: torch.Tensor, visibility: torch.Tensor, basis: Optional[torch.Tensor] = None):
    assert visibility.shape == indexes.shape, f"shape mismatch {visibility.shape} != {indexes.shape}"
    groups = [make_group(group, self.state) for group in self.param_groups]
    n = groups[0].num_points
    total_weight = get_total_weight(groups[0].state, n, device=visibility.device)
    running_vis = get_running_vis(groups[0].state, n, device=visibility.device)
    visibility = visibility.to(torch.float32).contiguous()
    weight = update_visibility(running_vis, visibility, indexes, total_weight, self.vis_beta)
    for group in groups:
      if group.grad is None:
        continue
      assert group.num_points == n, f"param shape {group.num_points} != {n}"
      apply_step(group, weight, indexes, total_weight, self.kernels, basis, grad_scale=visibility,
                 grad_smooth=self.vis_smooth)


class VisibilityAwareAdam(VisibilityOptimizer):
  def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-16, vis_beta=0.5, vis_smooth: float = 0.01,
               bias_correction=True, grad_clip: Optional[float] = None):
    super().__init__(ADAM, params, lr=lr, betas=betas, eps=eps, vis_beta=vis_beta, vis_smooth=vis_smooth,
                     bias_correction=bias_correction, grad_clip=grad_clip)


class VisibilityAwareLaProp(VisibilityOptimizer):
  def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-16, vis_beta=0.5, vis_smooth: float = 0.01,
               bias_correction=True, grad_clip: Optional[float] = None):
    super().__init__(LAPROP, params, lr=lr, betas=betas, eps=eps, vis_beta=vis_beta, vis_smooth=vis_smooth,
                     bias_correction=bias_correction, grad_clip=grad_clip)
