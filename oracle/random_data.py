"""Seeded synthetic inputs: a restatement of the reference's generators.  TEST INFRASTRUCTURE ONLY.

Follows taichi_splatting/tests/random_data.py:15-103 call for call (same torch CPU RNG
consumption order), so that `torch.manual_seed(s)` followed by these functions yields the
same clouds/cameras the reference's tests and benchmarks would see.
Containers are plain SimpleNamespace objects (tensordict is not in the image).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from . import torch_ops as T


def make_camera(T_camera_world, projection, image_size, near_plane, far_plane):
  fx, fy, cx, cy = [float(v) for v in projection]
  T_image_camera = torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=projection.dtype)
  T_ic4 = torch.eye(4, dtype=projection.dtype)
  T_ic4[0:3, 0:3] = T_image_camera
  return SimpleNamespace(T_camera_world=T_camera_world, projection=projection,
                         image_size=(int(image_size[0]), int(image_size[1])),
                         near_plane=float(near_plane), far_plane=float(far_plane),
                         depth_range=(float(near_plane), float(far_plane)),
                         T_image_camera=T_image_camera, T_image_world=T_ic4 @ T_camera_world,
                         camera_position=torch.inverse(T_camera_world)[0:3, 3])


def random_camera(pos_scale: float = 1., image_size: Optional[Tuple[int, int]] = None,
                  image_size_range=(256, 1024), near_plane=0.1):
  """tests/random_data.py:15-45"""
  q = F.normalize(torch.randn((1, 4)))
  t = torch.randn((3)) * pos_scale
  T_world_camera = T.join_rt(T.quat_to_mat(q), t)
  T_camera_world = torch.inverse(T_world_camera)
  if image_size is None:
    lo, hi = image_size_range
    image_size = [x.item() for x in torch.randint(size=(2,), low=lo, high=hi)]
  w, h = image_size
  cx, cy = torch.tensor([w / 2, h / 2]) + torch.randn(2) * (w / 20)
  fov = torch.deg2rad(torch.rand(1) * 70 + 30)
  fx = w / (2 * torch.tan(fov / 2))
  fy = h / (2 * torch.tan(fov / 2))
  projection = torch.tensor([fx, fy, cx, cy], dtype=torch.float32)
  return make_camera(T_camera_world, projection, (w, h), near_plane, near_plane * 1000.)


def fixed_camera(image_size: Tuple[int, int], fov_deg: float = 60.0, near_plane=0.1, far_plane=100.0,
                 yaw_deg: float = 0.0):
  """Benchmark camera: identity pose (optionally yawed), fov with the reference's
  convention fx = w / (2 tan(fov/2)), fy = h / (2 tan(fov/2)) (tests/random_data.py:33-35)."""
  w, h = image_size
  tan = math.tan(math.radians(fov_deg) / 2)
  projection = torch.tensor([w / (2 * tan), h / (2 * tan), w / 2, h / 2], dtype=torch.float32)
  Tcw = torch.eye(4)
  if yaw_deg != 0.0:
    a = math.radians(yaw_deg)
    Tcw[0, 0], Tcw[0, 2], Tcw[2, 0], Tcw[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
  return make_camera(Tcw, projection, (w, h), near_plane, far_plane)


def random_3d_gaussians(n, camera, scale_factor: float = 1.0, alpha_range=(0.1, 0.9), margin=0.0,
                        sh_degree: Optional[int] = None):
  """tests/random_data.py:48-75.  sh_degree=None -> feature (n,3) ~ U(0,1) as in the reference;
  otherwise feature (n,3,(deg+1)^2) with the SURVEY 8d builder's choice (DC (U-0.5)/0.2820948, rest
  N(0,0.1^2)), drawn AFTER all reference draws so the geometry matches the deg-0 cloud."""
  w, h = camera.image_size
  uv_pos = (torch.rand(n, 2) * (1 + margin) - margin * 0.5) * torch.tensor([w, h], dtype=torch.float32).unsqueeze(0)
  depth = T.inverse_ndc_depth(torch.rand(n), camera.near_plane * 2, camera.far_plane)
  position = T.unproject_points(uv_pos, depth.unsqueeze(1), camera.T_image_world)
  fx = camera.T_image_camera[0, 0]
  scale = (w / math.sqrt(n)) * (depth / fx) * scale_factor
  scaling = torch.randn(n, 3) * 0.5 + torch.log(scale).unsqueeze(1)
  rotation = F.normalize(torch.randn(n, 4), dim=1)
  low, high = alpha_range
  alpha = torch.rand(n) * (high - low) + low
  feature = torch.rand(n, 3)
  if sh_degree is not None:
    D = (sh_degree + 1)**2
    sh = torch.randn(n, 3, D) * 0.1
    sh[:, :, 0] = (feature - 0.5) / 0.282094791773878
    feature = sh
  return SimpleNamespace(position=position, log_scaling=scaling, rotation=rotation,
                         alpha_logit=torch.log(alpha / (1 - alpha)).unsqueeze(1), feature=feature)


def random_2d_gaussians(n, image_size: Tuple[int, int], num_channels=3, scale_factor=1.0,
                        alpha_range=(0.1, 0.9), depth_range=(0.0, 1.0)):
  """tests/random_data.py:78-103"""
  w, h = image_size
  position = torch.rand(n, 2) * torch.tensor([w, h], dtype=torch.float32).unsqueeze(0)
  depth = torch.rand((n, 1)) * (depth_range[1] - depth_range[0]) + depth_range[0]
  density_scale = scale_factor * w / (1 + math.sqrt(n))
  scaling = (torch.rand(n, 2) + 0.2) * density_scale
  rotation = torch.randn(n, 2)
  rotation = rotation / torch.norm(rotation, dim=1, keepdim=True)
  low, high = alpha_range
  alpha = torch.rand(n) * (high - low) + low
  return SimpleNamespace(position=position, depths=depth, log_scaling=torch.log(scaling),
                         rotation=rotation, alpha_logit=torch.log(alpha / (1 - alpha)),
                         feature=torch.rand(n, num_channels))


def packed_2d(g) -> torch.Tensor:
  """misc/renderer2d.py:16-33 applied to a random_2d_gaussians() cloud -> (N,7)."""
  return T.project_gaussians2d(g.position, g.log_scaling, g.rotation, g.alpha_logit)
