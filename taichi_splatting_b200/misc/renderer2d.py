"""2D Gaussian helpers used by tests, benchmarks and the image-fitting example
(reference: taichi_splatting/misc/renderer2d.py:16-33,134-150)."""
from numbers import Integral
from typing import Tuple

import torch

from ..data_types import Gaussians2D, RasterConfig
from ..rasterizer import rasterize


def project_gaussians2d(points: Gaussians2D) -> torch.Tensor:
  """Pure-torch "projection" of Gaussians2D to the packed (N,7) [mean, axis, sigma, alpha] form."""
  alpha = torch.sigmoid(points.alpha_logit)
  v1 = points.rotation / torch.norm(points.rotation, dim=1, keepdim=True)
  return torch.cat([points.position, v1, points.scaling, alpha.reshape(-1, 1)], dim=-1)


def point_basis(points: Gaussians2D, eps: float = 1e-4) -> torch.Tensor:
  """(N, 2, 2) basis of each Gaussian: its two axes as columns, scaled by sigma (reference :37-43).  The optimisers'
  `local_vector` groups step positions in this basis."""
  scale = torch.clamp_min(points.scaling, eps)
  v1 = points.rotation / torch.norm(points.rotation, dim=1, keepdim=True)
  v2 = torch.stack([-v1[..., 1], v1[..., 0]], dim=-1)
  return torch.stack([v1, v2], dim=2) * scale.unsqueeze(-2)


def render_gaussians(gaussians: Gaussians2D, image_size: Tuple[Integral, Integral],
                     raster_config: RasterConfig = RasterConfig()):
  gaussians2d = project_gaussians2d(gaussians)
  return rasterize(gaussians2d=gaussians2d, depth=torch.clamp(gaussians.depths, 0, 1),
                   features=gaussians.feature, image_size=image_size, config=raster_config)
