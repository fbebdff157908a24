"""Torch (CPU) restatement of the reference's differentiable point stages.  TEST INFRASTRUCTURE ONLY.

Restates, expression for expression, the reference's own pure-torch implementations
(which the reference uses as ITS test oracle for the Taichi kernels):

  project            <- torch_lib/projection.py:156-191 (apply), :21-41 (eig), :44-46 (ellipse_bounds),
                        :74-96 (project_with_jacobian), :57-71 (covariance_in_camera),
                        :98-106 (project_perspective_gaussian); torch_lib/transforms.py:5-51
  evaluate_sh_at     <- torch_lib/spherical_harmonics.py:17-43; basis constants
                        indexed_spherical_harmonics.py:38-106 (== torch_lib/rsh.py deg 0-3)
  ndc_depth          <- torch_lib/projection.py:120-123
  inverse_ndc_depth  <- torch_lib/projection.py:126-129
  project_gaussians2d<- misc/renderer2d.py:16-33

Pinned: tests/golden/*.npz hold outputs AND autograd gradients of the reference's torch_lib
(imported from /root/reference by tests/golden/make_golden.py); tests/test_oracle.py checks this
restatement against them.  Autograd through these functions is the gradient oracle, exactly as
tests/test_projection.py:86-91 and tests/test_spherical_harmonics.py:37-43 use torch_lib.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def quat_to_mat(quat: torch.Tensor) -> torch.Tensor:
  """torch_lib/transforms.py:5-15 (component order x,y,z,w: SURVEY D6)"""
  x, y, z, w = quat[..., 0], quat[..., 1], quat[..., 2], quat[..., 3]
  x2, y2, z2 = x * x, y * y, z * z
  return torch.stack([
      1 - 2 * y2 - 2 * z2, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y,
      2 * x * y + 2 * w * z, 1 - 2 * x2 - 2 * z2, 2 * y * z - 2 * w * x,
      2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x2 - 2 * y2
  ], dim=-1).reshape(quat.shape[:-1] + (3, 3))


def join_rt(r, t):
  T = torch.eye(4, device=r.device, dtype=r.dtype)
  T[0:3, 0:3] = r
  T[0:3, 3] = t
  return T


def make_homog(points):
  ones = torch.ones(points.shape[:-1] + (1,), dtype=points.dtype, device=points.device)
  return torch.cat([points, ones], dim=-1)


def transform44(transform, points):
  points = points.reshape([-1, 4, 1])
  return (transform.reshape([1, 4, 4]) @ points)[..., 0].reshape(-1, 4)


def eig(cov: torch.Tensor):
  """torch_lib/projection.py:21-41"""
  x, y, z = cov[..., 0, 0], cov[..., 0, 1], cov[..., 1, 1]
  tr = x + z
  det = x * z - y * y
  gap = tr**2 - 4 * det
  sqrt_gap = torch.sqrt(torch.clamp_min(gap, 0))
  lam1 = (tr + sqrt_gap) * 0.5
  lam2 = (tr - sqrt_gap) * 0.5
  v1 = F.normalize(torch.stack([x - lam2, y], -1), dim=-1)
  v2 = torch.stack([-v1[..., 1], v1[..., 0]], -1)
  return torch.stack([lam1, lam2], -1).sqrt(), v1, v2


def ndc_depth(depth: torch.Tensor, near: float, far: float) -> torch.Tensor:
  """torch_lib/projection.py:120-123 (eager torch semantics: scalars are rounded to the tensor dtype)"""
  return 1 - (1. / depth - 1. / far) / (1. / near - 1. / far)


def inverse_ndc_depth(ndc: torch.Tensor, near: float, far: float) -> torch.Tensor:
  """torch_lib/projection.py:126-129"""
  return 1.0 / ((1.0 - ndc) * (1 / near - 1 / far) + 1 / far)


def unproject_points(uv, depth, transform):
  """torch_lib/projection.py:60-64"""
  points = torch.cat([uv * depth, depth, torch.ones_like(depth)], dim=-1)
  transformed = transform44(torch.inverse(transform), points)
  return transformed[..., 0:3] / transformed[..., 3:4]


def project(position, log_scaling, rotation, alpha_logit, T_camera_world, projection, image_size,
            depth_range, blur_cov=0.0, clamp_margin=0.15, alpha_threshold=1. / 255., cull=True):
  """torch_lib/projection.py:156-191 -> points (V,7), depth (V,1), indexes (V,)"""
  point_in_camera = transform44(T_camera_world, make_homog(position))[:, :3]
  size = torch.tensor(image_size, dtype=position.dtype, device=position.device)

  f, c = projection[:2], projection[2:]
  z = point_in_camera[:, 2]
  uv = (point_in_camera[:, :2] * f) / z.unsqueeze(1) + c
  t = torch.clamp(uv, -clamp_margin * size, (1. + clamp_margin) * (size - 1))
  zero = torch.zeros_like(uv[:, 0])
  J = torch.stack([f[0] / z, zero, -(t[:, 0] - c[0]) / z,
                   zero, f[1] / z, -(t[:, 1] - c[1]) / z], dim=1).reshape(-1, 2, 3)

  W = T_camera_world[:3, :3]
  R = quat_to_mat(F.normalize(rotation, dim=-1))
  scale3 = log_scaling.exp()
  S = torch.eye(3, device=scale3.device, dtype=scale3.dtype).unsqueeze(0) * scale3.unsqueeze(1)
  m = W @ R @ S
  cov_cam = m @ m.transpose(1, 2)
  cov = torch.einsum('nij,njk,nkl->nil', J, cov_cam, J.transpose(1, 2))
  cov = cov + torch.eye(2, device=cov.device, dtype=cov.dtype) * blur_cov

  sigma, v1, v2 = eig(cov)
  alpha = alpha_logit.sigmoid()
  scale = sigma * torch.sqrt(2 * torch.log(alpha / alpha_threshold))
  ex1, ex2 = v1 * scale[:, 0:1], v2 * scale[:, 1:2]
  extent = torch.sqrt(ex1**2 + ex2**2)
  lower, upper = uv - extent, uv + extent

  in_view = ((z > depth_range[0]) & (z < depth_range[1]) & (upper > 0).all(1) &
             (lower < size.unsqueeze(0)).all(1))
  points = torch.cat([uv[:, :2], v1, sigma, alpha], dim=-1)
  if not cull:
    return points, z.unsqueeze(1), in_view
  vis_idx = in_view.nonzero(as_tuple=True)[0]
  return points[in_view], z[in_view].unsqueeze(1), vis_idx


def rsh_cart(degree: int, xyz: torch.Tensor) -> torch.Tensor:
  """Real SH basis, degree 0..3 (indexed_spherical_harmonics.py:38-106)."""
  x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
  out = [torch.full_like(x, 0.282094791773878)]
  if degree >= 1:
    out += [-0.48860251190292 * y, 0.48860251190292 * z, -0.48860251190292 * x]
  if degree >= 2:
    x2, y2, z2, xy, xz, yz = x * x, y * y, z * z, x * y, x * z, y * z
    out += [1.09254843059208 * xy, -1.09254843059208 * yz,
            0.94617469575756 * z2 - 0.31539156525252, -1.09254843059208 * xz,
            0.54627421529604 * x2 - 0.54627421529604 * y2]
  if degree >= 3:
    out += [-0.590043589926644 * y * (3.0 * x2 - y2), 2.89061144264055 * xy * z,
            0.304697199642977 * y * (1.5 - 7.5 * z2),
            1.24392110863372 * z * (1.5 * z2 - 0.5) - 0.497568443453487 * z,
            0.304697199642977 * x * (1.5 - 7.5 * z2), 1.44530572132028 * z * (x2 - y2),
            -0.590043589926644 * x * (x2 - 3.0 * y2)]
  return torch.stack(out, dim=-1)


def sh_degree(params: torch.Tensor) -> int:
  n = int(round(params.shape[2]**0.5))
  assert n * n == params.shape[2], f"SH feature count must be square, got {params.shape}"
  return n - 1


def evaluate_sh_at(params, points, indexes, camera_pos):
  """torch_lib/spherical_harmonics.py:31-43: clamp(sum_d Y_d(dir) * params[idx,c,d] + 0.5, 0, 1)"""
  dirs = F.normalize(points[indexes] - camera_pos.unsqueeze(0), dim=1)
  coeffs = rsh_cart(sh_degree(params), dirs)
  out = torch.einsum('nd,nkd->nk', coeffs, params[indexes])
  return torch.clamp(out + 0.5, 0., 1.)


def project_gaussians2d(position, log_scaling, rotation, alpha_logit):
  """misc/renderer2d.py:16-33 -> (N,7) [mean, axis, sigma, alpha]"""
  alpha = torch.sigmoid(alpha_logit)
  v1 = rotation / torch.norm(rotation, dim=1, keepdim=True)
  return torch.cat([position, v1, torch.exp(log_scaling), alpha.reshape(-1, 1)], dim=-1)
