"""Synthetic scenes for benchmarks and examples, in this package's own containers.

Statistically the reference's generators (taichi_splatting/tests/random_data.py:15-103; the benchmark
camera/cloud of SURVEY 8d): uv ~ U(image), depth = inverse-ndc(U(0,1), 2 near, far), log-scale
~ N(log((w / sqrt(n)) depth / fx * scale_factor), 0.5^2), unit random quaternions, alpha ~ U(lo, hi).
"""
import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from ..data_types import Gaussians2D, Gaussians3D
from ..perspective import CameraParams

SH_C0 = 0.282094791773878


def benchmark_camera(image_size: Tuple[int, int], fov_deg: float = 60.0, near_plane: float = 0.1,
                     far_plane: float = 100.0, yaw_deg: float = 0.0, device="cpu") -> CameraParams:
  w, h = image_size
  tan = math.tan(math.radians(fov_deg) / 2)
  projection = torch.tensor([w / (2 * tan), h / (2 * tan), w / 2, h / 2], dtype=torch.float32)
  T = torch.eye(4)
  if yaw_deg != 0.0:
    a = math.radians(yaw_deg)
    T[0, 0], T[0, 2], T[2, 0], T[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
  return CameraParams(projection=projection.to(device), T_camera_world=T.to(device), near_plane=near_plane,
                      far_plane=far_plane, image_size=(w, h))


def random_3d_gaussians(n: int, camera: CameraParams, scale_factor: float = 1.0, alpha_range=(0.1, 0.9),
                        margin: float = 0.0, sh_degree: Optional[int] = None, seed: int = 0) -> Gaussians3D:
  gen = torch.Generator().manual_seed(seed)
  w, h = camera.image_size
  cam = camera.to(device="cpu")
  uv = (torch.rand(n, 2, generator=gen) * (1 + margin) - margin * 0.5) * torch.tensor([w, h], dtype=torch.float32)
  near, far = cam.near_plane * 2, cam.far_plane
  depth = 1.0 / ((1.0 - torch.rand(n, generator=gen)) * (1 / near - 1 / far) + 1 / far)
  fx, fy, cx, cy = [float(v) for v in cam.projection]
  in_camera = torch.stack([(uv[:, 0] - cx) / fx * depth, (uv[:, 1] - cy) / fy * depth, depth, torch.ones(n)], dim=1)
  position = (torch.inverse(cam.T_camera_world) @ in_camera.T).T[:, :3].contiguous()
  scale = (w / math.sqrt(n)) * (depth / fx) * scale_factor
  log_scaling = torch.randn(n, 3, generator=gen) * 0.5 + torch.log(scale).unsqueeze(1)
  rotation = F.normalize(torch.randn(n, 4, generator=gen), dim=1)
  lo, hi = alpha_range
  alpha = torch.rand(n, generator=gen) * (hi - lo) + lo
  feature = torch.rand(n, 3, generator=gen)
  if sh_degree is not None:
    D = (sh_degree + 1)**2
    sh = torch.randn(n, 3, D, generator=gen) * 0.1   # most colours stay inside the clamp
    sh[:, :, 0] = (feature - 0.5) / SH_C0
    feature = sh
  return Gaussians3D(position=position, log_scaling=log_scaling, rotation=rotation,
                     alpha_logit=torch.log(alpha / (1 - alpha)).unsqueeze(1), feature=feature, batch_size=(n,))


def random_2d_gaussians(n: int, image_size: Tuple[int, int], num_channels: int = 3, scale_factor: float = 1.0,
                        alpha_range=(0.1, 0.9), depth_range=(0.0, 1.0), seed: int = 0) -> Gaussians2D:
  gen = torch.Generator().manual_seed(seed)
  w, h = image_size
  position = torch.rand(n, 2, generator=gen) * torch.tensor([w, h], dtype=torch.float32)
  depth = torch.rand((n, 1), generator=gen) * (depth_range[1] - depth_range[0]) + depth_range[0]
  scaling = (torch.rand(n, 2, generator=gen) + 0.2) * (scale_factor * w / (1 + math.sqrt(n)))
  rotation = F.normalize(torch.randn(n, 2, generator=gen), dim=1)
  lo, hi = alpha_range
  alpha = torch.rand(n, generator=gen) * (hi - lo) + lo
  return Gaussians2D(position=position, depths=depth, log_scaling=torch.log(scaling), rotation=rotation,
                     alpha_logit=torch.log(alpha / (1 - alpha)), feature=torch.rand(n, num_channels, generator=gen),
                     batch_size=(n,))
