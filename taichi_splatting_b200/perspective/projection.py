"""Perspective projection operator (reference: taichi_splatting/perspective/projection.py:27-255).

`apply` / `project_to_image` keep the reference's signatures and autograd contract; the device work is
gs_project_compact (single-pass project + cull + ordered compaction) / gs_project_bwd in libgsplat_b200.so
(csrc/projection.cu); gs_project_cull + gs_project_write remain as the two-kernel form of the same forward.
"""
from numbers import Integral
from beartype.typing import Tuple

import torch
from beartype import beartype

from .. import _lib
from ..data_types import Gaussians3D, RasterConfig
from .params import CameraParams


class _ProjectFunction(torch.autograd.Function):
  @staticmethod
  def forward(ctx, position, log_scaling, rotation, alpha_logit, T_camera_world, projection, image_size,
              depth_range, blur_cov, clamp_margin, alpha_threshold, want_ndc):
    _lib.require_cuda(position=position, log_scaling=log_scaling, rotation=rotation, alpha_logit=alpha_logit,
                      T_camera_world=T_camera_world, projection=projection)
    dtype, device = position.dtype, position.device
    sfx = _lib.suffix(dtype)
    n = position.shape[0]
    tensors = [t.detach().contiguous() for t in (position, log_scaling, rotation, alpha_logit,
                                                 T_camera_world, projection)]
    p = [_lib.ptr(t) for t in tensors]
    w, h = int(image_size[0]), int(image_size[1])
    near, far = float(depth_range[0]), float(depth_range[1])
    stream = _lib.stream_ptr(device)

    # single-pass project + cull + ordered compaction: outputs are written into capacity-n buffers, V arrives with the
    # one host sync of this operator (reference: torch.nonzero), and the results are the first V rows
    nbytes = _lib.c_size_t()
    _lib.call("gs_project_workspace_bytes", n, nbytes)
    ws = _lib.workspace(nbytes.value, device)
    word = _lib.host_word(device)
    points_n = torch.empty((n, 7), dtype=dtype, device=device)
    depth_n = torch.empty((n, 1), dtype=dtype, device=device)
    indexes_n = torch.empty((n,), dtype=torch.int64, device=device)
    ndc_n = torch.empty((n, 1), dtype=dtype, device=device) if want_ndc else None
    _lib.call(f"gs_project_compact_{sfx}", *p, n, w, h, near, far, float(blur_cov), float(clamp_margin),
              float(alpha_threshold), ws.data_ptr(), ws.numel(), _lib.ptr(points_n), _lib.ptr(depth_n),
              _lib.ptr(indexes_n), _lib.ptr(ndc_n), word.data_ptr(), stream)
    v = _lib.read_host_word(word, device)
    points, depth, indexes = points_n[:v], depth_n[:v], indexes_n[:v]
    ndc = ndc_n[:v] if want_ndc else None

    ctx.save_for_backward(*tensors, indexes)
    ctx.image_size, ctx.blur_cov, ctx.clamp_margin = (w, h), float(blur_cov), float(clamp_margin)
    ctx.mark_non_differentiable(indexes)
    if want_ndc:
      ctx.mark_non_differentiable(ndc)
      return points, depth, indexes, ndc
    return points, depth, indexes

  @staticmethod
  def backward(ctx, dpoints, ddepth, dindexes, *dndc):
    position, log_scaling, rotation, alpha_logit, T_camera_world, projection, indexes = ctx.saved_tensors
    device = position.device
    sfx = _lib.suffix(position.dtype)
    need = ctx.needs_input_grad
    grads = [torch.zeros_like(t) if need[i] else None
             for i, t in enumerate((position, log_scaling, rotation, alpha_logit, T_camera_world, projection))]
    if any(need[:6]) and indexes.shape[0] > 0:
      dpoints = dpoints.contiguous() if dpoints is not None else torch.zeros((indexes.shape[0], 7), dtype=position.dtype, device=device)
      ddepth = ddepth.contiguous() if ddepth is not None else torch.zeros((indexes.shape[0], 1), dtype=position.dtype, device=device)
      w, h = ctx.image_size
      _lib.call(f"gs_project_bwd_{sfx}", *[_lib.ptr(t) for t in (position, log_scaling, rotation, alpha_logit,
                                                                 T_camera_world, projection, indexes)],
                indexes.shape[0], w, h, ctx.blur_cov, ctx.clamp_margin, _lib.ptr(dpoints), _lib.ptr(ddepth),
                *[_lib.ptr(g) for g in grads], _lib.stream_ptr(device))
    return (*grads, None, None, None, None, None, None)


@beartype
def apply(position: torch.Tensor, log_scaling: torch.Tensor, rotation: torch.Tensor, alpha_logit: torch.Tensor,
          T_camera_world: torch.Tensor, projection: torch.Tensor, image_size: Tuple[Integral, Integral],
          depth_range: Tuple[float, float], blur_cov: float = 0.0, clamp_margin: float = 0.15,
          alpha_threshold: float = 1. / 255.) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
  """-> points (V,7) [mean, axis, sigma, alpha], depth (V,1) camera z, indexes (V,) int64 ascending."""
  return _ProjectFunction.apply(position, log_scaling, rotation, alpha_logit, T_camera_world, projection,
                                image_size, depth_range, blur_cov, clamp_margin, alpha_threshold, False)


def apply_with_ndc(position, log_scaling, rotation, alpha_logit, T_camera_world, projection, image_size,
                   depth_range, blur_cov=0.0, clamp_margin=0.15, alpha_threshold=1. / 255.):
  """Same as `apply`, plus ndc depth (V,1) computed in the same kernel (fuses torch_lib ndc_depth, R11)."""
  return _ProjectFunction.apply(position, log_scaling, rotation, alpha_logit, T_camera_world, projection,
                                image_size, depth_range, blur_cov, clamp_margin, alpha_threshold, True)


def camera_position_vjp(T_camera_world: torch.Tensor, camera_pos: torch.Tensor, d_camera_pos: torch.Tensor) -> torch.Tensor:
  """Gradient of the loss w.r.t. the (4,4) view matrix through camera_pos = inverse(T)[:3,3] (the reference gets it
  from autograd through torch.inverse, perspective/params.py:78-80).  With T = [A t; r s] evaluated at r = 0, s = 1:
  c = -A^-1 t, so dL/dt = u = -A^-T g, dL/dA = u c^T, dL/dr = -(g.c) c^T, dL/ds = -(g.c).  Tiny 3x3 work on the device,
  no host synchronisation (inv_ex skips the error read-back)."""
  T = T_camera_world.detach()
  g, c = d_camera_pos.to(T.dtype), camera_pos.detach().to(T.dtype)
  inv_A, _ = torch.linalg.inv_ex(T[:3, :3])
  u = -(inv_A.t() @ g)
  gc = torch.dot(g, c)
  d_T = torch.zeros_like(T)
  d_T[:3, :3] = torch.outer(u, c)
  d_T[:3, 3] = u
  d_T[3, :3] = -gc * c
  d_T[3, 3] = -gc
  return d_T


class _CameraPosition(torch.autograd.Function):
  @staticmethod
  def forward(ctx, T_camera_world):
    _lib.require_cuda(T_camera_world=T_camera_world)
    T = T_camera_world.detach().contiguous()
    out = torch.empty((3,), dtype=T.dtype, device=T.device)
    _lib.call(f"gs_camera_position_{_lib.suffix(T.dtype)}", _lib.ptr(T), _lib.ptr(out), _lib.stream_ptr(T.device))
    ctx.save_for_backward(T, out)
    return out

  @staticmethod
  def backward(ctx, d_out):
    T, out = ctx.saved_tensors
    return camera_position_vjp(T, out, d_out)


def camera_position(T_camera_world: torch.Tensor) -> torch.Tensor:
  """World-space camera centre of a (4,4) view matrix, computed on the device in one tiny kernel.  Same value and
  same gradient as CameraParams.camera_position (torch.inverse(T)[:3,3]) without the LU kernels and their host sync."""
  return _CameraPosition.apply(T_camera_world)


@beartype
def project_to_image(gaussians: Gaussians3D, camera_params: CameraParams, config: RasterConfig
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
  """Project 3D Gaussians to packed 2D Gaussians (EWA), culling to the view (reference :220-255)."""
  return apply(*gaussians.shape_tensors(), camera_params.T_camera_world, camera_params.projection,
               camera_params.image_size, camera_params.depth_range, config.blur_cov, config.clamp_margin,
               config.alpha_threshold)
