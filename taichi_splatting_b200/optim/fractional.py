"""Fractional / sparse Adam and LaProp (reference: taichi_splatting/optim/fractional.py:11-230).

Same classes, constructor arguments, `step(indexes, weight, basis=None)` signature, parameter-group keys and
state-dict layout as the reference.  Device work is one launch of gs_optim_step_f32 (csrc/optim.cu) per parameter
group: moments, bias correction, clip, mask_lr, point_lr, non-finite guard and the parameter update fused -- the
reference runs a Taichi kernel plus five to eight torch passes over (M, D) for the same result."""
import types
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from .. import _lib

ADAM, LAPROP = 0, 1


@dataclass
class Group:
  """One parameter group flattened to (N, D) (reference :11-57)."""
  name: str
  type: str
  param: torch.Tensor
  grad: Optional[torch.Tensor]
  state: dict
  lr: float
  betas: Tuple[float, float]
  eps: float
  bias_correction: bool
  clip: Optional[float]
  mask_lr: Optional[torch.Tensor]
  point_lr: Optional[torch.Tensor]

  @property
  def num_points(self):
    return self.param.shape[0]


def make_group(group, state) -> Group:
  n = len(group["params"])
  assert n == 1, f"expected 1 tensor in group {group['name']}, got {n}"
  params = group["params"][0]
  return Group(name=group["name"], type=group["type"], param=params.view(params.shape[0], -1),
               grad=params.grad.view(params.shape[0], -1) if params.grad is not None else None, state=state[params],
               lr=group["lr"], betas=group["betas"], eps=group["eps"], bias_correction=group["bias_correction"],
               clip=group.get("clip", None), mask_lr=group["mask_lr"], point_lr=group["point_lr"])


def get_total_weight(state: dict, n: int, device: torch.device):
  if 'total_weight' not in state:
    state['total_weight'] = torch.zeros(n, device=device, dtype=torch.float32)
  return state['total_weight']


def _moment_state(state: dict, param: torch.Tensor, vector: bool):
  """-> (first moment (N, D), second moment (N, D) | (N,)).  Key names and shapes follow the reference's
  optim/util.py:5-19 exactly -- including its swapped naming: the FIRST moment lives under 'v' -- so that state
  dicts are interchangeable."""
  if 'v' not in state:
    state['v'] = torch.zeros_like(param.view(param.shape[0], -1))
    state['m'] = (torch.zeros((param.shape[0],), dtype=param.dtype, device=param.device) if vector
                  else torch.zeros_like(param.view(param.shape[0], -1)))
  return state['v'], state['m']


def saturate(x: torch.Tensor):
  return 1 - 1 / torch.exp(2 * x)


def _launch(group: Group, algorithm: int, vector: bool, weight, indexes, total_weight, grad, grad_scale, grad_smooth,
            lr_step, param):
  first, second = _moment_state(group.state, group.param, vector)
  _lib.require_cuda(param=group.param, grad=grad, indexes=indexes, weight=weight)
  assert group.param.dtype == torch.float32 and grad.dtype == torch.float32, "optimisers are float32 (as upstream)"
  assert indexes.dtype == torch.int64, f"indexes must be int64, got {indexes.dtype}"
  assert group.param.is_contiguous(), "parameters must be contiguous"
  ptr = _lib.ptr
  mask_lr = group.mask_lr.to(torch.float32).contiguous().view(-1) if group.mask_lr is not None else None
  point_lr = group.point_lr.to(torch.float32).contiguous().view(-1) if group.point_lr is not None else None
  # contiguous copies (if any) stay referenced until the launch has been enqueued: a pointer taken from an unnamed
  # temporary would outlive its tensor
  indexes_c, weight_c, grad_c = indexes.contiguous(), weight.contiguous(), grad.contiguous()
  grad_scale_c = grad_scale.contiguous() if grad_scale is not None else None
  _lib.call("gs_optim_step_f32", algorithm, int(vector), int(group.bias_correction), ptr(indexes_c),
            ptr(weight_c), ptr(grad_scale_c), float(grad_smooth),
            indexes.shape[0], group.param.shape[1], ptr(first), ptr(second), ptr(total_weight), ptr(grad_c),
            float(group.lr), float(group.betas[0]), float(group.betas[1]), float(group.eps), ptr(lr_step),
            ptr(param) if param is not None else None, float(group.clip) if group.clip is not None else 0.0,
            ptr(mask_lr), ptr(point_lr), _lib.stream_ptr(group.param.device))


def weighted_step(group: Group, visible_weight: torch.Tensor, visible_indexes: torch.Tensor, total_weight: torch.Tensor,
                  algorithm: int, basis: Optional[torch.Tensor] = None, grad_scale: Optional[torch.Tensor] = None,
                  grad_smooth: float = 0.0):
  """The reference's weighted_step (:109-147): -> lr_step (M, D) after clip / basis / mask_lr / point_lr / finite guard.
  (`algorithm` replaces the reference's Taichi kernel module argument.)  Used for `local_vector` groups, whose basis
  change sits between the kernel and the masks; the other group types take the fused path in `apply_step`."""
  if group.type not in ("vector", "local_vector", "scalar"):
    raise ValueError(f"unknown group type {group.type}")
  vector = group.type in ("vector", "local_vector")
  grad = group.grad
  if group.type == "local_vector":
    assert basis is not None, "basis is required for local_vector optimizer"
    inv_basis = torch.linalg.inv(basis)
    grad = grad.clone()
    rows = grad[visible_indexes]
    if grad_scale is not None:
      rows = rows / (grad_scale.unsqueeze(1) + grad_smooth)
      grad_scale = None
    grad[visible_indexes] = torch.einsum('bij,bj->bi', inv_basis, rows)
  lr_step = group.param.new_zeros(visible_indexes.shape[0], group.param.shape[1])
  _launch(group, algorithm, vector, visible_weight, visible_indexes, total_weight, grad, grad_scale, grad_smooth,
          lr_step, None)
  if group.clip is not None:
    max_step = group.lr * group.clip
    lr_step.clamp_(-max_step, max_step)
  if group.type == "local_vector":
    lr_step = torch.einsum('bij,bj->bi', basis, lr_step)
  if group.mask_lr is not None:
    lr_step *= group.mask_lr.view(-1).unsqueeze(0)
  if group.point_lr is not None:
    lr_step *= group.point_lr[visible_indexes].unsqueeze(1)
  lr_step[~lr_step.isfinite()] = 0.0
  return lr_step


def apply_step(group: Group, weight, indexes, total_weight, algorithm: int, basis=None, grad_scale=None,
               grad_smooth: float = 0.0):
  """One optimiser step of a group on the visible rows: param[indexes] -= lr_step * saturate(weight) (:197-199)."""
  if group.type == "local_vector":
    lr_step = weighted_step(group, weight, indexes, total_weight, algorithm, basis, grad_scale, grad_smooth)
    group.param[indexes] -= lr_step * saturate(weight).unsqueeze(1)
    return
  if group.type not in ("vector", "scalar"):
    raise ValueError(f"unknown group type {group.type}")
  _launch(group, algorithm, group.type == "vector", weight, indexes, total_weight, group.grad, grad_scale, grad_smooth,
          None, group.param)


def resolve_algorithm(kernels) -> int:
  """ADAM / LAPROP; a reference-style kernel module (fractional_adam / fractional_laprop) is accepted too."""
  if isinstance(kernels, types.ModuleType):
    kernels = LAPROP if "laprop" in kernels.__name__ else ADAM
  assert kernels in (ADAM, LAPROP), f"unknown optimiser kernels {kernels!r}"
  return kernels


def group_defaults(lr, betas, eps, bias_correction, clip) -> dict:
  """Validated per-group defaults shared by every optimiser of this package (same checks and messages as upstream)."""
  for ok, what in ((lr > 0, f"learning rate: {lr}"), (eps > 0, f"epsilon: {eps}"),
                   (0.0 <= betas[0] < 1.0, f"beta1: {betas[0]}"), (0.0 <= betas[1] < 1.0, f"beta2: {betas[1]}")):
    assert ok, "Invalid " + what
  return dict(lr=lr, betas=betas, eps=eps, mask_lr=None, point_lr=None, type="scalar", bias_correction=bias_correction,
              clip=clip)


class FractionalOpt(torch.optim.Optimizer):
  """Reference :153-199.  `kernels` is ADAM or LAPROP (the reference passes the Taichi kernel module)."""

  def __init__(self, kernels, param_groups: list, lr=0.001, betas=(0.9, 0.999), eps=1e-16, bias_correction=True,
               clip: Optional[float] = None):
    self.kernels = resolve_algorithm(kernels)
    super().__init__(param_groups, group_defaults(lr, betas, eps, bias_correction, clip))

  @torch.no_grad()
  def step(self, indexes: torch.Tensor, weight: torch.Tensor, basis: Optional[torch.Tensor] = None):
    assert weight.shape == indexes.shape, f"shape mismatch {weight.shape} != {indexes.shape}"
    groups = [make_group(group, self.state) for group in self.param_groups]
    n = groups[0].param.shape[0]
    total_weight = get_total_weight(groups[0].state, n, device=weight.device)
    total_weight[indexes] += weight
    for group in groups:
      if group.grad is None:
        continue
      assert group.num_points == n, f"param shape {group.num_points} != {n}"
      apply_step(group, weight, indexes, total_weight, self.kernels, basis)


class FractionalAdam(FractionalOpt):
  def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-16, bias_correction=True):
    super().__init__(ADAM, params, lr, betas, eps, bias_correction)


class FractionalLaProp(FractionalOpt):
  def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-16, bias_correction=True):
    super().__init__(LAPROP, params, lr, betas, eps, bias_correction)


class SparseAdam(FractionalOpt):
  def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-16, bias_correction=True):
    super().__init__(ADAM, params, lr, betas, eps, bias_correction)

  def step(self, indexes: torch.Tensor, basis: Optional[torch.Tensor] = None):
    weight = torch.ones(indexes.shape[0], device=indexes.device, dtype=torch.float32)
    super().step(indexes, weight, basis)


class SparseLaProp(FractionalOpt):
  def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-16, bias_correction=True):
    super().__init__(LAPROP, params, lr, betas, eps, bias_correction)

  def step(self, indexes: torch.Tensor, basis: Optional[torch.Tensor] = None):
    weight = torch.ones(indexes.shape[0], device=indexes.device, dtype=torch.float32)
    super().step(indexes, weight, basis)
