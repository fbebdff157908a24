"""Render result containers (reference: taichi_splatting/rendering.py:27-160), tensordict-free."""
from dataclasses import dataclass, fields
from functools import cached_property
from typing import Any, Optional, Tuple

import torch

from .data_types import RasterConfig, _TensorFields
from .perspective.params import CameraParams


def ndc_depth(depth: torch.Tensor, near: float, far: float) -> torch.Tensor:
  """ndc from 0 (near) to 1 (far); reference torch_lib/projection.py:120-123."""
  return 1 - (1. / depth - 1. / far) / (1. / near - 1. / far)


def unpack(dc) -> dict:
  return {field.name: getattr(dc, field.name) for field in fields(dc)}


class RenderedPoints(_TensorFields):
  """Per in-view point outputs.  Fields as the reference TensorClass (rendering.py:27-101)."""
  _names = ("idx", "depths", "gaussians2d", "features")

  def __init__(self, idx, depths, gaussians2d, features, _prune_cost=None, _split_score=None,
               _visibility=None, attributes=None, batch_size=None):
    self._prune_cost, self._split_score, self._visibility = _prune_cost, _split_score, _visibility
    self.attributes = attributes
    super().__init__(idx=idx, depths=depths, gaussians2d=gaussians2d, features=features, batch_size=batch_size)

  def apply(self, fn, batch_size=None):
    opt = lambda t: None if t is None else fn(t)
    return RenderedPoints(fn(self.idx), fn(self.depths), fn(self.gaussians2d), fn(self.features),
                          opt(self._prune_cost), opt(self._split_score), opt(self._visibility), self.attributes)

  @property
  def prune_cost(self):
    assert self._prune_cost is not None, "No prune cost information available (render with config.compute_point_heuristic=True)"
    return self._prune_cost

  @property
  def split_score(self):
    assert self._split_score is not None, "No split score information available (render with config.compute_point_heuristic=True)"
    return self._split_score

  @property
  def visibility(self):
    assert self._visibility is not None, "No visibility information available (render with config.compute_visibility=True)"
    return self._visibility

  @property
  def screen_scale(self):
    return self.gaussians2d[:, 4:6]

  @property
  def opacity(self):
    return self.gaussians2d[:, 6]

  @property
  def visible_mask(self) -> torch.Tensor:
    return self.visibility > 0.0

  @cached_property
  def visible(self) -> 'RenderedPoints':
    return self[self.visible_mask]

  @property
  def num_visible(self) -> int:
    return int(self.visible_mask.sum().item())

  def full_mask(self, n: int) -> torch.Tensor:
    mask = torch.zeros((n,), dtype=torch.bool, device=self.idx.device)
    mask[self.idx] = self.visible_mask
    return mask

  def full_visibility(self, n: int) -> torch.Tensor:
    vis = torch.zeros((n,), dtype=self.visibility.dtype, device=self.visibility.device)
    vis[self.idx] = self.visibility
    return vis

  def gaussian_scale(self, alpha_threshold: float = 1.0 / 255):
    return torch.sqrt(2 * torch.log(self.opacity / alpha_threshold))


@dataclass(frozen=True, kw_only=True)
class Rendering:
  """Renderer outputs (reference rendering.py:104-160)."""
  image: torch.Tensor                                  # (H, W, C)
  image_weight: torch.Tensor                           # (H, W)
  depth_image: Optional[torch.Tensor] = None           # always None in the reference (D15)
  median_depth_image: Optional[torch.Tensor] = None    # (H, W)
  points: RenderedPoints
  camera: CameraParams
  config: RasterConfig
  glo_feature: Optional[torch.Tensor] = None

  @cached_property
  def ndc_image(self) -> torch.Tensor:
    return ndc_depth(self.depth_image, self.camera.near_plane, self.camera.far_plane)

  @cached_property
  def median_ndc_image(self) -> torch.Tensor:
    return ndc_depth(self.median_depth_image, self.camera.near_plane, self.camera.far_plane)

  @property
  def visible_idx(self) -> torch.Tensor:
    return self.points.idx[self.points.visible_mask]

  @property
  def in_view_idx(self) -> torch.Tensor:
    return self.points.idx

  @property
  def visible_points(self) -> RenderedPoints:
    return self.points[self.points.visible_mask]

  @property
  def image_size(self) -> Tuple[int, int]:
    return self.camera.image_size

  def detach(self):
    return Rendering(**{k: x.detach() if hasattr(x, 'detach') else x for k, x in unpack(self).items()})
