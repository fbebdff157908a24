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


# ---- densification helpers (reference: misc/renderer2d.py:46-131; host-side torch there as well) ---------------------
def gaussian_covariance(points: Gaussians2D) -> torch.Tensor:
  """(N, 2, 2) covariance B B^T of each Gaussian, B = point_basis."""
  basis = point_basis(points)
  return basis @ basis.transpose(1, 2)


def sample_gaussians(points: Gaussians2D) -> torch.Tensor:
  """One offset per Gaussian drawn from its own distribution: B z, z ~ N(0, I)."""
  z = torch.randn_like(points.position)
  return torch.einsum("nij,nj->ni", point_basis(points), z)


def repeat_sample_gaussians(samples: torch.Tensor, points: Gaussians2D, n: int = 2) -> torch.Tensor:
  """samples (N, n, 2) in each Gaussian's own basis -> offsets (N, n, 2) in image space."""
  return torch.einsum("nij,nkj->nki", point_basis(points), samples.view(-1, n, 2))


def split_with_offsets(points: Gaussians2D, offsets: torch.Tensor, depth_noise: float = 1e-2) -> Gaussians2D:
  """Every Gaussian replaced by `n` copies displaced by offsets (N, n, 2); depths jittered so that the copies keep a
  definite order (and stay positive)."""
  num_points, n, _ = offsets.shape
  copies = points.apply(lambda t: torch.repeat_interleave(t, repeats=n, dim=0), batch_size=(num_points * n,))
  depths = torch.clamp_min(copies.depths + torch.randn_like(copies.depths) * depth_noise, 1e-6)
  return copies.replace(position=copies.position + offsets.reshape(-1, 2), depths=depths)


def split_gaussians2d(points: Gaussians2D, n: int = 2, scaling: float = None, depth_noise: float = 1e-2) -> Gaussians2D:
  """The splitting step of 3D Gaussian splatting in 2D: `n` children per Gaussian, sampled from (half the spread of)
  the parent and shrunk by `scaling` (default 1 / sqrt(n))."""
  import math
  samples = 0.5 * torch.randn((points.batch_size[0], n, 2), device=points.position.device)
  offsets = repeat_sample_gaussians(samples, points, n)
  shrink = math.log(scaling if scaling is not None else 1 / math.sqrt(n))
  return split_with_offsets(points.replace(log_scaling=points.log_scaling + shrink), offsets, depth_noise)


def uniform_split_gaussians2d(points: Gaussians2D, n: int = 2, scaling: float = None, depth_noise: float = 1e-2,
                              sep: float = 0.7, random_axis: bool = False, eps: float = 1e-6) -> Gaussians2D:
  """Deterministic split along ONE axis of each Gaussian -- its longest, or (random_axis) one drawn with probability
  proportional to the axis lengths: `n` children evenly spaced in [-sep, sep] sigma along that axis, which alone is
  shrunk by `scaling` (default 1 / sqrt(n))."""
  import math
  if random_axis:
    probs = torch.nn.functional.normalize(points.scaling + eps, p=1, dim=1)
    axis = torch.multinomial(probs, num_samples=1).squeeze(1)
  else:
    axis = torch.argmax(points.log_scaling, dim=1)
  along = torch.nn.functional.one_hot(axis, num_classes=2).to(points.position.dtype)      # (N, 2)
  steps = torch.linspace(-sep, sep, n, device=points.position.device)
  samples = steps.view(1, n, 1) * along.view(-1, 1, 2)
  offsets = repeat_sample_gaussians(samples, points, n)
  shrink = scaling if scaling is not None else math.sqrt(n) / n
  shrunk = points.set_scaling(points.scaling * (along * shrink + (1 - along)))
  return split_with_offsets(shrunk, offsets, depth_noise)
