"""Value types crossing the operator boundary (reference: taichi_splatting/data_types.py:16-145).

Same names and fields as the reference so callers can switch packages.  tensordict is not a dependency
here: Gaussians3D / Gaussians2D are small field containers with the TensorClass methods the render path
and its callers use (`to`, `replace`, indexing, `batch_size`, `shape_tensors`, `packed`, ...).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, fields, replace
from typing import List, Tuple

import torch


@dataclass(frozen=True, eq=True, kw_only=True)
class RasterConfig:
  """Hashable raster configuration (reference data_types.py:16-46), plus `forward_saturate_eps`."""
  tile_size: int = 16
  pixel_stride: Tuple[int, int] = (2, 2)   # reference backward thread tiling; accepted, not needed here
  clamp_margin: float = 0.15
  antialias: bool = False
  blur_cov: float = 0.3
  clamp_max_alpha: float = 0.99
  alpha_threshold: float = 1. / 255.
  saturate_threshold: float = 0.9999
  use_alpha_blending: bool = True
  compute_point_heuristic: bool = False
  compute_visibility: bool = False
  median_threshold: float = 0.25
  # Not in the reference (its forward never stops early): opt-in early-out -- the forward pass stops
  # compositing a pixel once its remaining transmittance is <= this value, bounding the image error by
  # eps * max|feature| (and truncating per-point visibility by the same weights).  The default 0 reproduces the
  # reference exactly (drop-in parity).
  forward_saturate_eps: float = 0.0

  def __post_init__(self):
    assert self.tile_size in (8, 16, 32), f"tile_size {self.tile_size} not supported (8, 16, 32)"
    assert len(self.pixel_stride) == 2


class _TensorFields:
  """Minimal stand-in for tensordict.TensorClass: named tensor fields sharing a leading batch dim."""
  _names: Tuple[str, ...] = ()

  def __init__(self, batch_size=None, **kwargs):
    missing = [n for n in self._names if n not in kwargs]
    assert not missing, f"{type(self).__name__}: missing fields {missing}"
    for n in self._names:
      setattr(self, n, kwargs[n])
    n0 = getattr(self, self._names[0]).shape[0]
    self.batch_size = torch.Size(batch_size) if batch_size is not None else torch.Size((n0,))
    self.__post_init__()

  def __post_init__(self):
    pass

  def to_dict(self):
    return {n: getattr(self, n) for n in self._names}

  def apply(self, fn, batch_size=None):
    return type(self)(**{n: fn(t) for n, t in self.to_dict().items()}, batch_size=batch_size)

  def replace(self, **kwargs):
    d = self.to_dict()
    d.update(kwargs)
    return type(self)(**d)

  def to(self, *args, **kwargs):
    def conv(t):
      if t.is_floating_point():
        return t.to(*args, **kwargs)
      kw = {k: v for k, v in kwargs.items() if k != "dtype"}
      a = [x for x in args if not isinstance(x, torch.dtype)]
      return t.to(*a, **kw)
    return self.apply(conv)

  def cuda(self):
    return self.apply(lambda t: t.cuda())

  def cpu(self):
    return self.apply(lambda t: t.cpu())

  def detach(self):
    return self.apply(torch.Tensor.detach)

  def clone(self):
    return self.apply(torch.Tensor.clone)

  def requires_grad_(self, requires_grad: bool = True):
    for t in self.to_dict().values():
      if t.is_floating_point():
        t.requires_grad_(requires_grad)
    return self

  def __getitem__(self, idx):
    return self.apply(lambda t: t[idx])

  def __len__(self):
    return int(self.batch_size[0])

  @property
  def device(self):
    return getattr(self, self._names[0]).device

  def __repr__(self):
    inner = ", ".join(f"{n}={tuple(getattr(self, n).shape)}" for n in self._names)
    return f"{type(self).__name__}({inner})"


class Gaussians3D(_TensorFields):
  """position (N,3), log_scaling (N,3), rotation (N,4) quaternion xyzw, alpha_logit (N,1),
  feature (N,C) or (N,3,(deg+1)^2).  Reference: data_types.py:57-114."""
  _names = ("position", "log_scaling", "rotation", "alpha_logit", "feature")

  def __post_init__(self):
    assert self.position.shape[1] == 3, f"Expected shape (N, 3), got {self.position.shape}"
    assert self.log_scaling.shape[1] == 3, f"Expected shape (N, 3), got {self.log_scaling.shape}"
    assert self.rotation.shape[1] == 4, f"Expected shape (N, 4), got {self.rotation.shape}"
    assert self.alpha_logit.shape[1] == 1, f"Expected shape (N, 1), got {self.alpha_logit.shape}"

  def packed(self):
    return torch.cat([self.position, self.log_scaling, self.rotation, self.alpha_logit], dim=-1)

  def shape_tensors(self):
    return (self.position, self.log_scaling, self.rotation, self.alpha_logit)

  def scaled(self, scale: float) -> 'Gaussians3D':
    return self.replace(position=self.position * scale, log_scaling=math.log(scale) + self.log_scaling)

  def translated(self, translation: torch.Tensor) -> 'Gaussians3D':
    return self.replace(position=self.position + translation.view(1, 3))

  @property
  def scale(self):
    return torch.exp(self.log_scaling)

  @property
  def alpha(self):
    return torch.sigmoid(self.alpha_logit)

  @staticmethod
  def concat_batch(gaussians: List['Gaussians3D']) -> 'Gaussians3D':
    keys = gaussians[0].to_dict().keys()
    return Gaussians3D(**{k: torch.cat([getattr(g, k) for g in gaussians], dim=0) for k in keys})


def inverse_sigmoid(x: torch.Tensor):
  return torch.log(x / (1 - x))


class Gaussians2D(_TensorFields):
  """position (N,2), depths (N,1), log_scaling (N,2), rotation (N,2), alpha_logit (N,), feature (N,C).
  Reference: data_types.py:122-145."""
  _names = ("position", "depths", "log_scaling", "rotation", "alpha_logit", "feature")

  @property
  def opacity(self):
    return self.alpha_logit.sigmoid()

  @property
  def scaling(self):
    return torch.exp(self.log_scaling)

  def set_scaling(self, scaling) -> 'Gaussians2D':
    return self.replace(log_scaling=torch.log(scaling))


__all__ = ["RasterConfig", "Gaussians3D", "Gaussians2D", "inverse_sigmoid", "replace", "fields"]
