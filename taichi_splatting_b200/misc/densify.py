"""Heuristics-driven split / prune of a Gaussian cloud (N4; reference: examples/fit_image_gaussians.py:169-230).

The raster backward accumulates two numbers per Gaussian (rasterizer/backward.py:190-194, exposed as
`Rendering.points.prune_cost` / `.split_score` and `RasterOut.point_heuristic`): sum (alpha dL/dalpha)^2 -- how much the
loss depends on the Gaussian -- and sum |alpha dL/dalpha dpdf/dmean|_1 -- how hard the loss pulls on its position.
`find_split_prune` turns them into masks (prune the least useful, split the most pulled-on so that the cloud reaches a
target size); `split_prune` applies them to a Gaussians2D cloud and carries a sparse optimiser's per-point state along
(kept rows keep theirs, children start from zero).  The reference holds parameters and optimiser state in its
`ParameterClass` container; here the optimiser is any torch optimiser whose state tensors have one row per point
(every optimiser of `taichi_splatting_b200.optim`)."""
from typing import Dict, Optional, Sequence, Tuple

import torch

from ..data_types import Gaussians2D
from .renderer2d import uniform_split_gaussians2d


def take_n(t: torch.Tensor, n: int, descending: bool = False) -> torch.Tensor:
  """Mask of the n largest (descending) or smallest entries of a 1-D tensor."""
  mask = torch.zeros_like(t, dtype=torch.bool)
  if n > 0:
    mask[torch.argsort(t, descending=descending)[:n]] = True
  return mask


def randomize_n(t: torch.Tensor, n: int) -> torch.Tensor:
  """Mask of n entries drawn without replacement with probability proportional to t."""
  mask = torch.zeros_like(t, dtype=torch.bool)
  if n > 0:
    mask[torch.multinomial(torch.nn.functional.normalize(t, dim=0), n, replacement=False)] = True
  return mask


def find_split_prune(n: int, target: int, n_prune: int, prune_cost: torch.Tensor, densify_score: torch.Tensor
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
  """-> (split_mask, prune_mask), disjoint.  The n_prune cheapest points go; enough of the highest-scoring points are
  split (each into two) to bring the cloud from n to `target`.  A point chosen for both is left alone."""
  prune_mask = take_n(prune_cost, n_prune, descending=False)
  pruned = int(prune_mask.sum().item())
  split_mask = take_n(densify_score, max(0, (target - n) + pruned), descending=True)
  both = split_mask & prune_mask
  return split_mask ^ both, prune_mask ^ both


def resize_optimizer_state(optimizer: torch.optim.Optimizer, old_params: Sequence[torch.Tensor],
                           new_params: Sequence[torch.Tensor], keep: torch.Tensor, num_new: int) -> None:
  """Re-points every single-tensor parameter group from old_params[i] to new_params[i] (rows = kept rows of the old
  tensor followed by num_new new rows).  State tensors with one row per point follow their rows; new rows are zero."""
  n_old = old_params[0].shape[0]
  mapping = {id(o): nw for o, nw in zip(old_params, new_params)}
  for group in optimizer.param_groups:
    for i, p in enumerate(group["params"]):
      if id(p) not in mapping:
        continue
      new_p = mapping[id(p)]
      state = optimizer.state.pop(p, None)
      if state:
        resized = {}
        for key, value in state.items():
          if torch.is_tensor(value) and value.ndim >= 1 and value.shape[0] == n_old:
            pad = value.new_zeros((num_new, *value.shape[1:]))
            resized[key] = torch.cat([value[keep], pad], dim=0)
          else:
            resized[key] = value
        optimizer.state[new_p] = resized
      group["params"][i] = new_p


def split_prune(gaussians: Gaussians2D, t: float, target: int, prune_rate: float,
                heuristics: Tuple[torch.Tensor, torch.Tensor], optimizer: Optional[torch.optim.Optimizer] = None,
                random_axis: bool = True) -> Tuple[Gaussians2D, Dict[str, int]]:
  """One densification step at training progress t in [0, 1]: prune prune_rate * n * (1 - t) points by prune cost,
  split the top split scores towards `target` points.  Returns the new cloud (leaf tensors, requires_grad as before)
  and the counts; `optimizer`'s groups and per-point state are moved onto the new tensors."""
  prune_cost, split_score = heuristics
  n = gaussians.batch_size[0]
  split_mask, prune_mask = find_split_prune(n=n, target=target, n_prune=int(prune_rate * n * (1 - t)),
                                            prune_cost=prune_cost, densify_score=split_score)
  with torch.no_grad():
    children = uniform_split_gaussians2d(gaussians[split_mask].detach(), random_axis=random_axis)
    keep = ~(split_mask | prune_mask)
    kept = gaussians[keep].detach()
    names = gaussians.to_dict().keys()
    merged = {k: torch.cat([getattr(kept, k), getattr(children, k)], dim=0).contiguous() for k in names}
  new = Gaussians2D(**merged)
  old_tensors = [getattr(gaussians, k) for k in names]
  for k, old in zip(names, old_tensors):
    getattr(new, k).requires_grad_(old.requires_grad)
  if optimizer is not None:
    resize_optimizer_state(optimizer, old_tensors, [getattr(new, k) for k in names], keep, children.batch_size[0])
  return new, dict(split=int(split_mask.sum().item()), prune=int(prune_mask.sum().item()))
