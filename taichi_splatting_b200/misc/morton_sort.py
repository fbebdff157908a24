"""Morton (Z-order) sorting of point clouds (reference: taichi_splatting/misc/morton_sort.py:95-130).

argsort(points, resolution): 63-bit Morton codes on a 2^21 grid of cell size `resolution` anchored at the cloud's
minimum corner (gs_morton_codes64), then a stable radix argsort (gs_sort_pairs) -- the reference's
code_points64_kernel + cuda_lib.radix_argsort.  Ordering the cloud this way keeps Gaussians that are close in space
close in memory, which makes the rasteriser's per-tile record gathers local."""
import ctypes

import torch

from .. import _lib


def grid_at_resolution(points: torch.Tensor, resolution: float, size: int = 2**20):
  """-> (lower (3,), upper (3,), size): reference :95-99."""
  lower = points.min(dim=0).values
  upper = lower + size * resolution
  return lower, upper, size


def morton_codes(points: torch.Tensor, resolution: float, size: int = 2**20):
  """-> (codes (N,) int64 holding the unsigned 63-bit codes, ids (N,) int32 = arange)."""
  assert points.ndim == 2 and points.shape[1] == 3, f"points must be (N, 3), got {points.shape}"
  _lib.require_cuda(points=points)
  p = points.detach().to(torch.float32).contiguous()
  n = p.shape[0]
  codes = torch.empty((n,), dtype=torch.int64, device=p.device)
  ids = torch.empty((n,), dtype=torch.int32, device=p.device)
  if n == 0:
    return codes, ids
  lower, upper, size = grid_at_resolution(p, resolution, size)
  inc = (upper - lower) / float(size)     # Grid.get_inc (:44-46), float32 like the Taichi struct fields
  f3 = ctypes.c_float * 3
  _lib.call("gs_morton_codes64", _lib.ptr(p), n, f3(*lower.tolist()), f3(*inc.tolist()), size, _lib.ptr(codes),
            _lib.ptr(ids), _lib.stream_ptr(p.device))
  return codes, ids


def argsort(points: torch.Tensor, resolution: float) -> torch.Tensor:
  """Indexes (N,) int32 that put `points` in Morton order (reference :119-125)."""
  codes, ids = morton_codes(points, resolution, size=2**20)
  n = codes.shape[0]
  if n == 0:
    return ids
  nbytes = _lib.c_size_t()
  _lib.call("gs_sort_pairs_workspace_bytes", n, 8, nbytes)
  ws = _lib.workspace(nbytes.value, points.device)
  codes_out, ids_out = torch.empty_like(codes), torch.empty_like(ids)
  _lib.call("gs_sort_pairs", _lib.ptr(codes), _lib.ptr(ids), _lib.ptr(codes_out), _lib.ptr(ids_out), n, 8, 0, 63,
            ws.data_ptr(), ws.numel(), _lib.stream_ptr(points.device))
  return ids_out


def sort(points: torch.Tensor, resolution: float) -> torch.Tensor:
  return points[argsort(points, resolution).long()]
