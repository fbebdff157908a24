"""Renderer facade (reference: taichi_splatting/renderer.py:22-121): project -> SH | gather -> map_to_tiles
-> rasterize (-> median-depth raster), each stage one of this package's operators."""
from dataclasses import replace

import torch
from beartype import beartype

from . import _lib
from .data_types import Gaussians3D, RasterConfig
from .mapper.tile_mapper import bin_and_sort, map_to_tiles
from .perspective import CameraParams
from .perspective.projection import apply_with_ndc, camera_position
from .rasterizer.function import (fused_median_supported, make_digest, rasterize_with_tiles,
                                  rasterize_with_tiles_and_median, tuned_supported)
from .rendering import RenderedPoints, Rendering, ndc_depth
from .spherical_harmonics import check_sh_degree, evaluate_sh_at


class _RenderFunction(torch.autograd.Function):
  """The whole render path as ONE autograd node.

  Same stages, same kernels and same results as chaining the operators (project -> SH | gather -> map_to_tiles ->
  rasterize), but the host enqueues the front end without per-operator autograd / validation overhead, so the GPU
  does not idle between the short front-end kernels, and the backward is three launches issued back to back."""

  @staticmethod
  def forward(ctx, position, log_scaling, rotation, alpha_logit, feature, T_camera_world, projection, camera, config,
              use_sh, use_depth16, render_median_depth, sh_exchange=None):
    _lib.require_cuda(position=position, log_scaling=log_scaling, rotation=rotation, alpha_logit=alpha_logit,
                      feature=feature, T_camera_world=T_camera_world, projection=projection)
    dtype, device = position.dtype, position.device
    sfx = _lib.suffix(dtype)
    call, ptr = _lib.call, _lib.ptr
    stream = _lib.stream_ptr(device)
    n = position.shape[0]
    w, h = int(camera.image_size[0]), int(camera.image_size[1])
    near, far = float(camera.near_plane), float(camera.far_plane)
    blur, margin, thr = float(config.blur_cov), float(config.clamp_margin), float(config.alpha_threshold)
    tensors = [t.detach().contiguous() for t in (position, log_scaling, rotation, alpha_logit, T_camera_world, projection)]
    feature_c = feature.detach().contiguous()
    pin = [ptr(t) for t in tensors]

    # ---- projection: cull -> V -> compacted write (+ ndc depth) ----
    nbytes = _lib.c_size_t()
    call("gs_project_workspace_bytes", n, nbytes)
    ws_proj = _lib.workspace(nbytes.value, device)
    word = _lib.host_word(device)
    call(f"gs_project_cull_{sfx}", *pin, n, w, h, near, far, blur, margin, thr, ws_proj.data_ptr(), ws_proj.numel(),
         word.data_ptr(), stream)
    cam_pos = None
    if use_sh:   # independent of V: enqueue while the host waits for it
      cam_pos = torch.empty((3,), dtype=dtype, device=device)
      call(f"gs_camera_position_{sfx}", pin[4], ptr(cam_pos), stream)
    v = _lib.read_host_word(word, device)

    g2d = torch.empty((v, 7), dtype=dtype, device=device)
    depths = torch.empty((v, 1), dtype=dtype, device=device)
    ndc = torch.empty((v, 1), dtype=dtype, device=device)
    indexes = torch.empty((v,), dtype=torch.int64, device=device)
    call(f"gs_project_write_{sfx}", *pin, n, w, h, near, far, blur, margin, ws_proj.data_ptr(), ptr(g2d), ptr(depths),
         ptr(indexes), ptr(ndc), stream)

    # ---- features: SH at the visible set, or a plain gather ----
    if use_sh:
      degree = check_sh_degree(feature_c)
      channels = feature_c.shape[1]
      features = torch.empty((v, channels), dtype=dtype, device=device)
      call(f"gs_sh_fwd_{sfx}", ptr(feature_c), pin[0], ptr(indexes), ptr(cam_pos), v, channels, degree, ptr(features), stream)
    else:
      assert feature_c.ndim == 2, f"Features must be (N, C) if use_sh=False, got {feature_c.shape}"
      features = feature_c[indexes]
    F = features.shape[1]

    # ---- tile mapper (fp32 only, like the reference): two-level ordering, one host read (K) ----
    g32 = g2d if dtype == torch.float32 else g2d.float()
    d32 = (ndc if dtype == torch.float32 else ndc.float()).view(-1)
    overlap_to_point, tile_ranges, _, _, _ = bin_and_sort(g32, d32, (w, h), config, use_depth16)
    k = overlap_to_point.shape[0]
    ranges = tile_ranges.view(-1, 2)

    # ---- rasteriser (+ fused median depth) ----
    image = torch.empty((h, w, F), dtype=dtype, device=device)
    alpha = torch.empty((h, w), dtype=dtype, device=device)
    heuristic = (torch.zeros((v, 2), dtype=dtype, device=device) if config.compute_point_heuristic
                 else torch.empty((0, 2), dtype=dtype, device=device))
    visibility = (torch.zeros((v,), dtype=dtype, device=device) if config.compute_visibility
                  else torch.empty((0,), dtype=dtype, device=device))
    vis_ptr = ptr(visibility) if config.compute_visibility else None
    cfg = _lib.raster_config_c(config)
    median = torch.empty((0,), dtype=dtype, device=device)
    digest = torch.empty((0, 16), dtype=torch.float32, device=device)
    fused_median = render_median_depth and fused_median_supported(config, F, dtype)
    if tuned_supported(config, F, dtype) and (fused_median or not render_median_depth):
      # raster records, written once and gathered by the forward and the backward kernel
      digest = make_digest(g2d, features, depths.view(-1) if fused_median else None, config)
      if fused_median:
        median = torch.empty((h, w), dtype=dtype, device=device)
      call("gs_raster_fwd_digest_f32", ptr(digest), ptr(ranges), ptr(overlap_to_point), v, k, w, h, F, cfg,
           float(config.median_threshold), ptr(image), ptr(alpha), vis_ptr, ptr(median) if fused_median else None,
           stream)
    else:
      call(f"gs_raster_fwd_{sfx}", ptr(g2d), ptr(features), ptr(ranges), ptr(overlap_to_point), v, k, w, h, F, cfg,
           ptr(image), ptr(alpha), vis_ptr, stream)
      if render_median_depth:   # the reference's second, non-blending pass (renderer.py:77-82)
        dcfg = _lib.raster_config_c(replace(config, use_alpha_blending=False, saturate_threshold=config.median_threshold,
                                            compute_visibility=False, compute_point_heuristic=False))
        median3 = torch.empty((h, w, 1), dtype=dtype, device=device)
        call(f"gs_raster_fwd_{sfx}", ptr(g2d), ptr(depths), ptr(ranges), ptr(overlap_to_point), v, k, w, h, 1, dcfg,
             ptr(median3), ptr(torch.empty((h, w), dtype=dtype, device=device)), None, stream)
        median = median3.squeeze(-1)

    ctx.save_for_backward(*tensors, feature_c, indexes, g2d, features, image, overlap_to_point, ranges,
                          cam_pos if cam_pos is not None else torch.empty(0, device=device), digest)
    ctx.meta = (config, (w, h), blur, margin, bool(use_sh), heuristic)
    ctx.sh_exchange = sh_exchange
    ctx.set_materialize_grads(False)
    ctx.mark_non_differentiable(alpha, indexes, visibility, heuristic, median, overlap_to_point, tile_ranges)
    return image, alpha, g2d, depths, indexes, features, visibility, heuristic, median, overlap_to_point, tile_ranges

  @staticmethod
  def backward(ctx, d_image, d_alpha, d_g2d, d_depths, d_indexes, d_features, *unused):
    (position, log_scaling, rotation, alpha_logit, T_camera_world, projection, feature, indexes, g2d, features, image,
     overlap_to_point, ranges, cam_pos, digest) = ctx.saved_tensors
    config, (w, h), blur, margin, use_sh, heuristic = ctx.meta
    dtype, device = position.dtype, position.device
    sfx = _lib.suffix(dtype)
    call, ptr = _lib.call, _lib.ptr
    stream = _lib.stream_ptr(device)
    v, F = g2d.shape[0], features.shape[1]
    need = ctx.needs_input_grad
    need_geom = any(need[i] for i in (0, 1, 2, 3, 5, 6))

    # ---- rasteriser backward: gradients of the packed 2D Gaussians and per-point features ----
    grad_g = d_g2d.clone() if d_g2d is not None else torch.zeros_like(g2d)
    grad_f = d_features.clone() if d_features is not None else torch.zeros_like(features)
    if d_image is not None and v > 0:
      out_ptrs = (ptr(grad_g) if need_geom else None, ptr(grad_f) if need[4] else None,
                  ptr(heuristic) if config.compute_point_heuristic else None)
      if digest.shape[0] == v:
        call("gs_raster_bwd_digest_f32", ptr(digest), ptr(ranges), ptr(overlap_to_point), ptr(image),
             ptr(d_image.contiguous()), v, overlap_to_point.shape[0], w, h, F, _lib.raster_config_c(config),
             *out_ptrs, stream)
      else:
        call(f"gs_raster_bwd_{sfx}", ptr(g2d), ptr(features), ptr(ranges), ptr(overlap_to_point), ptr(image),
             ptr(d_image.contiguous()), v, overlap_to_point.shape[0], w, h, F, _lib.raster_config_c(config),
             *out_ptrs, stream)

    # ---- features (part 1): view-parallel runs launch the exchange of the SH-gradient factors now, so that the
    # all-gather overlaps the projection backward ----
    d_feature = None
    exchange = ctx.sh_exchange if (need[4] and use_sh and dtype == torch.float32) else None
    if exchange is not None and exchange.world <= 1:
      exchange = None
    pending = exchange.start(feature, indexes, features, grad_f, cam_pos) if exchange is not None else None

    # ---- projection backward ----
    grads = [torch.zeros_like(t) if need[i] else None
             for t, i in ((position, 0), (log_scaling, 1), (rotation, 2), (alpha_logit, 3), (T_camera_world, 5), (projection, 6))]
    if need_geom and v > 0:
      dd = d_depths.contiguous() if d_depths is not None else torch.zeros((v, 1), dtype=dtype, device=device)
      call(f"gs_project_bwd_{sfx}", ptr(position), ptr(log_scaling), ptr(rotation), ptr(alpha_logit), ptr(T_camera_world),
           ptr(projection), ptr(indexes), v, w, h, blur, margin, ptr(grad_g), ptr(dd), *[ptr(g) for g in grads], stream)

    # ---- features (part 2) ----
    if exchange is not None:
      d_feature = exchange.finish(pending, feature, position, check_sh_degree(feature))
    elif need[4]:
      all_rows_written = use_sh and v == feature.shape[0]
      d_feature = torch.empty_like(feature) if all_rows_written else torch.zeros_like(feature)
      if v > 0:
        if use_sh:
          call(f"gs_sh_bwd_{sfx}", ptr(feature), ptr(position), ptr(indexes), ptr(cam_pos), ptr(grad_f), ptr(features), v,
               feature.shape[1], check_sh_degree(feature), 1, ptr(d_feature), None, None, stream)
        else:
          d_feature.index_copy_(0, indexes, grad_f)
    return (grads[0], grads[1], grads[2], grads[3], d_feature, grads[4], grads[5], None, None, None, None, None, None)


@beartype
def render_gaussians(gaussians: Gaussians3D, camera_params: CameraParams, config: RasterConfig = RasterConfig(),
                     use_sh: bool = False, render_depth: bool = False, use_depth16: bool = False,
                     render_median_depth: bool = False) -> Rendering:
  """Complete renderer for 3D Gaussians; same parameters and result type as the reference (renderer.py:22-59).
  `render_depth` is accepted and unused, as in the reference.  Runs as one fused autograd node
  (`_RenderFunction`); `render_projected` below is the operator-by-operator composition of the same stages."""
  outs = _RenderFunction.apply(*gaussians.shape_tensors(), gaussians.feature, camera_params.T_camera_world,
                               camera_params.projection, camera_params, config, use_sh, use_depth16,
                               render_median_depth)
  return _wrap_rendering(outs, camera_params, config, render_median_depth)


def _wrap_rendering(outs, camera_params, config, render_median_depth) -> Rendering:
  (image, alpha, g2d, depths, indexes, features, visibility, heuristic, median, _, _) = outs
  points = RenderedPoints(
      idx=indexes, depths=depths, gaussians2d=g2d,
      _visibility=visibility if config.compute_visibility else None,
      _prune_cost=heuristic[:, 0] if config.compute_point_heuristic else None,
      _split_score=heuristic[:, 1] if config.compute_point_heuristic else None,
      features=features, attributes=None, batch_size=(depths.shape[0],))
  return Rendering(image=image, image_weight=alpha, depth_image=None,
                   median_depth_image=median if render_median_depth else None, points=points, camera=camera_params,
                   config=config)


def render_gaussians_unfused(gaussians: Gaussians3D, camera_params: CameraParams, config: RasterConfig = RasterConfig(),
                             use_sh: bool = False, use_depth16: bool = False,
                             render_median_depth: bool = False) -> Rendering:
  """The reference's own composition (renderer.py:50-59): one autograd node per operator."""
  gaussians2d, depths, indexes, ndc = apply_with_ndc(
      *gaussians.shape_tensors(), camera_params.T_camera_world, camera_params.projection,
      camera_params.image_size, camera_params.depth_range, config.blur_cov, config.clamp_margin,
      config.alpha_threshold)
  if use_sh:
    features = evaluate_sh_at(gaussians.feature, gaussians.position.detach(), indexes,
                              camera_position(camera_params.T_camera_world), unique_indexes=True)
  else:
    features = gaussians.feature[indexes]
    assert len(features.shape) == 2, f"Features must be (N, C) if use_sh=False, got {features.shape}"
  return render_projected(indexes, gaussians2d, features, depths, camera_params, config,
                          use_depth16=use_depth16, render_median_depth=render_median_depth, ndc_depths=ndc)


def render_projected(indexes: torch.Tensor, gaussians2d: torch.Tensor, features: torch.Tensor,
                     depths: torch.Tensor, camera_params: CameraParams, config: RasterConfig,
                     use_depth16: bool = False, render_median_depth: bool = False, ndc_depths=None) -> Rendering:
  if ndc_depths is None:
    ndc_depths = ndc_depth(depths.detach(), camera_params.near_plane, camera_params.far_plane)

  overlap_to_point, tile_overlap_ranges = map_to_tiles(gaussians2d, ndc_depths, image_size=camera_params.image_size,
                                                       config=config, use_depth16=use_depth16)
  ranges = tile_overlap_ranges.view(-1, 2)
  median_depth = None
  if render_median_depth:
    # the reference runs a second, non-blending raster pass over features=depths (:77-82); here the same value
    # comes out of the main pass (no gradient flows through it: the reference's non-blending backward is
    # disabled upstream)
    raster, median_depth = rasterize_with_tiles_and_median(
        gaussians2d, features, depths, overlap_to_point=overlap_to_point, tile_overlap_ranges=ranges,
        image_size=camera_params.image_size, config=config)
  else:
    raster = rasterize_with_tiles(gaussians2d, features, tile_overlap_ranges=ranges,
                                  overlap_to_point=overlap_to_point, image_size=camera_params.image_size, config=config)

  points = RenderedPoints(
      idx=indexes, depths=depths, gaussians2d=gaussians2d,
      _visibility=raster.visibility if config.compute_visibility else None,
      _prune_cost=raster.point_heuristic[:, 0] if config.compute_point_heuristic else None,
      _split_score=raster.point_heuristic[:, 1] if config.compute_point_heuristic else None,
      features=features, attributes=None, batch_size=(depths.shape[0],))

  return Rendering(image=raster.image, image_weight=raster.image_weight, depth_image=None,
                   median_depth_image=median_depth, points=points, camera=camera_params, config=config)


def viewspace_gradient(gaussians2d: torch.Tensor):
  assert gaussians2d.shape[1] == 7, f"Expected packed 2D gaussians (N,7), got {gaussians2d.shape}"
  assert gaussians2d.grad is not None, "Expected gradients on gaussians2d, run backward first with gaussians2d.retain_grad()"
  return torch.norm(gaussians2d.grad[:, :2], dim=1)
