"""Spherical-harmonics colour operator (reference: taichi_splatting/indexed_spherical_harmonics.py:23-188,
spherical_harmonics.py:139-170).  Device work: gs_sh_fwd / gs_sh_bwd (csrc/sh.cu)."""
import math

import torch
from beartype import beartype

from . import _lib


def check_sh_degree(sh_features: torch.Tensor) -> int:
  assert len(sh_features.shape) == 3, f"SH features must have 3 dimensions, got {sh_features.shape}"
  n_sh = sh_features.shape[2]
  n = int(math.sqrt(n_sh))
  assert n * n == n_sh, f"SH feature count must be square, got {n_sh} ({sh_features.shape})"
  assert 0 <= n - 1 <= 3, f"SH degree must be between 0 and 3, got {n - 1}"
  return n - 1


class _ShFunction(torch.autograd.Function):
  @staticmethod
  def forward(ctx, params, points, indexes, camera_pos, unique_indexes):
    _lib.require_cuda(sh_params=params, positions=points, indexes=indexes, camera_pos=camera_pos)
    sfx = _lib.suffix(params.dtype)
    device = params.device
    degree = check_sh_degree(params)
    params, points, camera_pos = (t.detach().contiguous() for t in (params, points, camera_pos))
    indexes = indexes.contiguous()
    assert indexes.dtype == torch.int64, f"indexes must be int64, got {indexes.dtype}"
    v, channels = indexes.shape[0], params.shape[1]
    out = torch.empty((v, channels), dtype=params.dtype, device=device)
    _lib.call(f"gs_sh_fwd_{sfx}", _lib.ptr(params), _lib.ptr(points), _lib.ptr(indexes), _lib.ptr(camera_pos),
              v, channels, degree, _lib.ptr(out), _lib.stream_ptr(device))
    ctx.save_for_backward(params, points, indexes, camera_pos, out)
    ctx.degree, ctx.unique = degree, bool(unique_indexes)
    ctx.mark_non_differentiable(indexes)
    return out

  @staticmethod
  def backward(ctx, doutput):
    params, points, indexes, camera_pos, out = ctx.saved_tensors
    sfx = _lib.suffix(params.dtype)
    need = ctx.needs_input_grad
    d_params = torch.zeros_like(params) if need[0] else None
    d_points = torch.zeros_like(points) if need[1] else None
    d_cam = torch.zeros_like(camera_pos) if need[3] else None
    if (need[0] or need[1] or need[3]) and indexes.shape[0] > 0:
      doutput_c = doutput.contiguous()   # named: the pointer must not outlive a temporary copy
      _lib.call(f"gs_sh_bwd_{sfx}", _lib.ptr(params), _lib.ptr(points), _lib.ptr(indexes), _lib.ptr(camera_pos),
                _lib.ptr(doutput_c), _lib.ptr(out), indexes.shape[0], params.shape[1], ctx.degree, int(ctx.unique),
                _lib.ptr(d_params), _lib.ptr(d_points), _lib.ptr(d_cam), _lib.stream_ptr(params.device))
    return d_params, d_points, None, d_cam, None


@beartype
def evaluate_sh_at(sh_params: torch.Tensor,   # (M, C, (degree+1)^2)
                   positions: torch.Tensor,   # (M, 3)
                   indexes: torch.Tensor,     # (V,) int64 into 0..M
                   camera_pos: torch.Tensor,  # (3,)
                   unique_indexes: bool = False) -> torch.Tensor:   # (V, C)
  """clamp(SH(normalize(p[idx] - cam)) . params[idx] + 0.5, 0, 1).  `unique_indexes=True` (an addition to
  the reference signature) promises no repeated index and makes the backward write plain stores."""
  return _ShFunction.apply(sh_params, positions, indexes, camera_pos, unique_indexes)


@beartype
def evaluate_sh(sh_params: torch.Tensor, positions: torch.Tensor, camera_pos: torch.Tensor) -> torch.Tensor:
  """Non-indexed variant (reference spherical_harmonics.py:159-170).  The reference kernel omits the
  +0.5 / clamp and returns a wrong backward arity; this one is evaluate_sh_at over all points."""
  idx = torch.arange(positions.shape[0], device=positions.device, dtype=torch.int64)
  return _ShFunction.apply(sh_params, positions, idx, camera_pos, True)


__all__ = ["evaluate_sh", "evaluate_sh_at", "check_sh_degree"]
