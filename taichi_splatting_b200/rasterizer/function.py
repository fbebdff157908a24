"""Rasteriser operator (reference: taichi_splatting/rasterizer/function.py:19-165).

Same signatures / autograd contract as the reference; device work is gs_raster_fwd / gs_raster_bwd
(csrc/raster_fwd.cu, raster_bwd.cu, raster_generic.cu)."""
from numbers import Integral
from beartype.typing import NamedTuple, Optional, Tuple

import torch
from beartype import beartype

from .. import _lib
from ..data_types import RasterConfig
from ..mapper.tile_mapper import map_to_tiles

RasterOut = NamedTuple('RasterOut', [
    ('image', torch.Tensor),
    ('image_weight', torch.Tensor),
    ('point_heuristic', Optional[torch.Tensor]),
    ('visibility', Optional[torch.Tensor]),
])


def tuned_supported(config: RasterConfig, num_features: int, dtype) -> bool:
  """Configurations of the tuned kernels (raster_fwd.cu / raster_bwd_t.cu), which read the raster digest."""
  return dtype == torch.float32 and config.tile_size == 16 and not config.antialias and 1 <= num_features <= 4


def fused_median_supported(config: RasterConfig, num_features: int, dtype) -> bool:
  """The fused median-depth output exists in the tuned, alpha-blending kernel only."""
  return tuned_supported(config, num_features, dtype) and config.use_alpha_blending


def make_digest(gaussians2d: torch.Tensor, features: torch.Tensor, depths: Optional[torch.Tensor],
                config: RasterConfig) -> torch.Tensor:
  """(V,16) fp32 raster records (gs_raster_digest_f32): written once, gathered by forward and backward."""
  v = gaussians2d.shape[0]
  digest = torch.empty((v, 16), dtype=torch.float32, device=gaussians2d.device)
  if v > 0:
    _lib.call("gs_raster_digest_f32", _lib.ptr(gaussians2d), _lib.ptr(features),
              _lib.ptr(depths) if depths is not None else None, v, features.shape[1], _lib.raster_config_c(config),
              _lib.ptr(digest), _lib.stream_ptr(gaussians2d.device))
  return digest


def pack_records(digest: torch.Tensor, ranges: torch.Tensor, overlap_to_point: torch.Tensor, image_size,
                 num_features: int):
  """Per-overlap raster records in sorted order (gs_raster_pack_f32): (K, 12 | 16) sweep records and (K, 4) flush
  records.  Written once per frame after the sort; the forward and backward kernels fetch a tile's batch from them
  with one bulk copy (TMA engine) instead of gathering through overlap_to_point."""
  k = overlap_to_point.shape[0]
  device = digest.device
  records = torch.empty((k, 12 if num_features <= 3 else 16), dtype=torch.float32, device=device)
  flush_records = torch.empty((k, 4), dtype=torch.float32, device=device)
  if k > 0:
    _lib.call("gs_raster_pack_f32", _lib.ptr(digest), _lib.ptr(ranges), _lib.ptr(overlap_to_point), k,
              int(image_size[0]), int(image_size[1]), num_features, _lib.ptr(records), _lib.ptr(flush_records),
              _lib.stream_ptr(device))
  return records, flush_records


def rasterize_with_tiles_and_median(gaussians2d, features, depths, overlap_to_point, tile_overlap_ranges, image_size,
                                    config):
  """rasterize_with_tiles plus the reference's median-depth pass (renderer.py:77-82) -> (RasterOut, median (H,W)).
  One fused kernel when supported, else the reference's two passes."""
  if fused_median_supported(config, features.shape[1], gaussians2d.dtype):
    *out, median = _RasterFunction.apply(gaussians2d, features, overlap_to_point, tile_overlap_ranges, image_size,
                                         config, depths)
    return RasterOut(*out), median
  from dataclasses import replace
  raster = rasterize_with_tiles(gaussians2d, features, overlap_to_point, tile_overlap_ranges, image_size, config)
  depth_config = replace(config, use_alpha_blending=False, saturate_threshold=config.median_threshold,
                         compute_visibility=False, compute_point_heuristic=False)
  raster_depth = rasterize_with_tiles(gaussians2d.detach(), depths.detach(), overlap_to_point, tile_overlap_ranges,
                                      image_size, depth_config)
  return raster, raster_depth.image.squeeze(-1)


class _RasterFunction(torch.autograd.Function):
  @staticmethod
  def forward(ctx, gaussians, features, overlap_to_point, tile_overlap_ranges, image_size, config,
              median_depths=None):
    _lib.require_cuda(gaussians2d=gaussians, features=features, overlap_to_point=overlap_to_point,
                      tile_overlap_ranges=tile_overlap_ranges)
    dtype, device = gaussians.dtype, gaussians.device
    sfx = _lib.suffix(dtype)
    assert features.dtype == dtype, f"features dtype {features.dtype} != gaussians dtype {dtype}"
    assert overlap_to_point.dtype == torch.int32 and tile_overlap_ranges.dtype == torch.int32
    w, h = int(image_size[0]), int(image_size[1])
    ts = config.tile_size
    n_tiles = ((w + ts - 1) // ts) * ((h + ts - 1) // ts)
    ranges = tile_overlap_ranges.contiguous().view(-1, 2)
    assert ranges.shape[0] == n_tiles, f"tile_overlap_ranges has {ranges.shape[0]} tiles, image needs {n_tiles}"
    g = gaussians.detach().contiguous()
    f = features.detach().contiguous()
    o2p = overlap_to_point.contiguous()
    v, F = g.shape[0], f.shape[1]

    image = torch.empty((h, w, F), dtype=dtype, device=device)
    alpha = torch.empty((h, w), dtype=dtype, device=device)
    heuristic = (torch.zeros((v, 2), dtype=dtype, device=device) if config.compute_point_heuristic
                 else torch.empty((0, 2), dtype=dtype, device=device))
    visibility = (torch.zeros((v,), dtype=dtype, device=device) if config.compute_visibility
                  else torch.empty((0,), dtype=dtype, device=device))
    cfg = _lib.raster_config_c(config)
    vis_ptr = _lib.ptr(visibility) if config.compute_visibility else None
    median = None
    digest = torch.empty((0, 16), dtype=torch.float32, device=device)
    packed = None
    if median_depths is not None:
      assert fused_median_supported(config, F, dtype), "fused median depth: unsupported configuration"
      median = torch.empty((h, w), dtype=dtype, device=device)
    if tuned_supported(config, F, dtype):
      d = median_depths.detach().contiguous().view(-1) if median_depths is not None else None
      digest = make_digest(g, f, d, config)
      if config.use_alpha_blending:
        packed = pack_records(digest, ranges, o2p, (w, h), F)
        _lib.call("gs_raster_fwd_packed_f32", _lib.ptr(packed[0]), _lib.ptr(ranges), _lib.ptr(o2p), v, o2p.shape[0], w,
                  h, F, cfg, float(config.median_threshold), _lib.ptr(image), _lib.ptr(alpha), vis_ptr,
                  _lib.ptr(median) if median is not None else None, _lib.stream_ptr(device))
      else:
        _lib.call("gs_raster_fwd_digest_f32", _lib.ptr(digest), _lib.ptr(ranges), _lib.ptr(o2p), v, o2p.shape[0], w, h,
                  F, cfg, float(config.median_threshold), _lib.ptr(image), _lib.ptr(alpha), vis_ptr,
                  _lib.ptr(median) if median is not None else None, _lib.stream_ptr(device))
    else:
      _lib.call(f"gs_raster_fwd_{sfx}", _lib.ptr(g), _lib.ptr(f), _lib.ptr(ranges), _lib.ptr(o2p), v, o2p.shape[0],
                w, h, F, cfg, _lib.ptr(image), _lib.ptr(alpha), vis_ptr, _lib.stream_ptr(device))

    ctx.config, ctx.image_size = config, (w, h)
    ctx.heuristic = heuristic
    ctx.packed = packed
    ctx.save_for_backward(g, f, image, o2p, ranges, digest)
    ctx.mark_non_differentiable(alpha, heuristic, visibility)
    if median is not None:
      ctx.mark_non_differentiable(median)
      return image, alpha, heuristic, visibility, median
    return image, alpha, heuristic, visibility

  @staticmethod
  def backward(ctx, grad_image, grad_alpha, grad_heuristic, grad_visibility, *grad_median):
    g, f, image, o2p, ranges, digest = ctx.saved_tensors
    config, (w, h) = ctx.config, ctx.image_size
    sfx = _lib.suffix(g.dtype)
    need_g, need_f = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    grad_g = torch.zeros_like(g) if need_g else None
    grad_f = torch.zeros_like(f) if need_f else None
    if need_g or need_f or config.compute_point_heuristic:
      grad_image_c = grad_image.contiguous()   # named: the pointer must not outlive a temporary copy
      cfg = _lib.raster_config_c(config)
      heur_ptr = _lib.ptr(ctx.heuristic) if config.compute_point_heuristic else None
      if ctx.packed is not None:
        _lib.call("gs_raster_bwd_packed_f32", _lib.ptr(ctx.packed[0]), _lib.ptr(ctx.packed[1]), _lib.ptr(ranges),
                  _lib.ptr(o2p), _lib.ptr(image), _lib.ptr(grad_image_c), None, g.shape[0], o2p.shape[0], w,
                  h, f.shape[1], cfg, _lib.ptr(grad_g), _lib.ptr(grad_f), heur_ptr, _lib.stream_ptr(g.device))
      elif tuned_supported(config, f.shape[1], g.dtype):
        _lib.call("gs_raster_bwd_digest_f32", _lib.ptr(digest), _lib.ptr(ranges), _lib.ptr(o2p), _lib.ptr(image),
                  _lib.ptr(grad_image_c), g.shape[0], o2p.shape[0], w, h, f.shape[1], cfg,
                  _lib.ptr(grad_g), _lib.ptr(grad_f), heur_ptr, _lib.stream_ptr(g.device))
      else:
        _lib.call(f"gs_raster_bwd_{sfx}", _lib.ptr(g), _lib.ptr(f), _lib.ptr(ranges), _lib.ptr(o2p), _lib.ptr(image),
                  _lib.ptr(grad_image_c), g.shape[0], o2p.shape[0], w, h, f.shape[1], cfg,
                  _lib.ptr(grad_g), _lib.ptr(grad_f), heur_ptr, _lib.stream_ptr(g.device))
    return grad_g, grad_f, None, None, None, None, None


@beartype
def rasterize_with_tiles(gaussians2d: torch.Tensor, features: torch.Tensor, overlap_to_point: torch.Tensor,
                         tile_overlap_ranges: torch.Tensor, image_size: Tuple[Integral, Integral],
                         config: RasterConfig) -> RasterOut:
  """Rasterise packed 2D Gaussians (N,7) with features (N,F) given tile overlap information.
  Returns RasterOut(image (H,W,F), image_weight (H,W), point_heuristic (N,2), visibility (N,))."""
  assert gaussians2d.ndim == 2 and gaussians2d.shape[1] == 7, f"gaussians2d must be Nx7, got {gaussians2d.shape}"
  assert features.ndim == 2 and features.shape[0] == gaussians2d.shape[0], \
      f"Size mismatch: got {gaussians2d.shape}, {features.shape}"
  return RasterOut(*_RasterFunction.apply(gaussians2d, features, overlap_to_point, tile_overlap_ranges,
                                          image_size, config))


def rasterize(gaussians2d: torch.Tensor, depth: torch.Tensor, features: torch.Tensor,
              image_size: Tuple[Integral, Integral], config: RasterConfig, use_depth16: bool = False) -> RasterOut:
  """map_to_tiles + rasterize_with_tiles (reference :133-165)."""
  assert gaussians2d.shape[0] == depth.shape[0] == features.shape[0], \
      f"Size mismatch: got {gaussians2d.shape}, {depth.shape}, {features.shape}"
  overlap_to_point, tile_overlap_ranges = map_to_tiles(gaussians2d, depth, image_size=image_size, config=config,
                                                       use_depth16=use_depth16)
  return rasterize_with_tiles(gaussians2d, features, tile_overlap_ranges=tile_overlap_ranges.view(-1, 2),
                              overlap_to_point=overlap_to_point, image_size=image_size, config=config)
