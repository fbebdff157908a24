"""Renderer facade (reference: taichi_splatting/renderer.py:22-121): project -> SH | gather -> map_to_tiles
-> rasterize (-> median-depth raster), each stage one of this package's operators."""
import os
from dataclasses import replace

import torch
from beartype import beartype

from . import _lib
from .data_types import Gaussians3D, RasterConfig
from .mapper.tile_mapper import ORDERING, bin_and_sort, bin_and_sort_binned, map_to_tiles
from .perspective import CameraParams
from .perspective.projection import apply_with_ndc, camera_position, camera_position_vjp
from .rasterizer.function import (fused_median_supported, pack_records, rasterize_with_tiles,
                                  rasterize_with_tiles_and_median, tuned_supported)
from .rendering import RenderedPoints, Rendering, ndc_depth
from .spherical_harmonics import check_sh_degree, evaluate_sh_at


# GS_FUSED_HOST=0 chains the per-stage entry points from Python instead of calling the whole-frame drivers
# (csrc/render.cu): same kernels and results, more host time per launch (A/B switch; also what bench.py uses for
# its per-stage timings)
_FUSED_HOST = os.environ.get("GS_FUSED_HOST", "1") != "0"
_ws_sizes = {}
_k_capacity = {}   # device index -> capacity (in overlaps) to give the K-sized buffers of the next frame


def _workspace_sizes(n: int):
  """Workspace bytes of the projection, depth-order and scan stages for n Gaussians (cached)."""
  if n not in _ws_sizes:
    out = []
    for name in ("gs_project_workspace_bytes", "gs_depth_order_workspace_bytes", "gs_tile_scan_workspace_bytes"):
      nbytes = _lib.c_size_t()
      _lib.call(name, n, nbytes)
      out.append(int(nbytes.value))
    _ws_sizes[n] = tuple(out)
  return _ws_sizes[n]


def _event_handle(ev):
  return ev.cuda_event if ev is not None else None


# Optional lists of (start, end) torch.cuda.Event pairs, one pair consumed per frame, which the whole-frame drivers
# record around the raster launches on the caller's stream (bench.py times the dominant kernel with them).  The
# events must be created with enable_timing=True and recorded once beforehand so that their handles exist.
raster_events = {"fwd": None, "bwd": None}


def _next_event_pair(which):
  pairs = raster_events[which]
  return pairs.pop(0) if pairs else None


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
    channels = feature.shape[1] if feature.ndim >= 2 else 0
    ctx.fused_host = (_FUSED_HOST and dtype == torch.float32 and config.use_alpha_blending
                      and tuned_supported(config, channels, dtype) and feature.ndim == (3 if use_sh else 2))
    assert getattr(sh_exchange, "kind", None) != "tile" or ctx.fused_host, \
        "tile-sharded rendering runs through the whole-frame drivers (fp32, tile 16, alpha blending, GS_FUSED_HOST != 0)"
    if ctx.fused_host:
      return _RenderFunction._forward_fused_host(ctx, position, log_scaling, rotation, alpha_logit, feature,
                                                 T_camera_world, projection, camera, config, use_sh, use_depth16,
                                                 render_median_depth, sh_exchange)
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

    # ---- projection: single-pass project + cull + ordered compaction (+ ndc depth) into capacity-n buffers -> V ----
    nbytes = _lib.c_size_t()
    call("gs_project_workspace_bytes", n, nbytes)
    ws_proj = _lib.workspace(nbytes.value, device)
    word = _lib.host_word(device)
    g2d_n = torch.empty((n, 7), dtype=dtype, device=device)
    depths_n = torch.empty((n, 1), dtype=dtype, device=device)
    ndc_n = torch.empty((n, 1), dtype=dtype, device=device)
    indexes_n = torch.empty((n,), dtype=torch.int64, device=device)
    call(f"gs_project_compact_{sfx}", *pin, n, w, h, near, far, blur, margin, thr, ws_proj.data_ptr(), ws_proj.numel(),
         ptr(g2d_n), ptr(depths_n), ptr(indexes_n), ptr(ndc_n), word.data_ptr(), stream)
    cam_pos = None
    if use_sh:   # independent of V: enqueue while the host waits for it
      cam_pos = torch.empty((3,), dtype=dtype, device=device)
      call(f"gs_camera_position_{sfx}", pin[4], ptr(cam_pos), stream)
    v = _lib.read_host_word(word, device)
    g2d, depths, ndc, indexes = g2d_n[:v], depths_n[:v], ndc_n[:v], indexes_n[:v]

    # ---- features: SH at the visible set, or a plain gather ----
    if use_sh:
      degree = check_sh_degree(feature_c)
      channels = feature_c.shape[1]
      features = torch.empty((v, channels), dtype=dtype, device=device)
      call(f"gs_sh_fwd_{sfx}", ptr(feature_c), pin[0], ptr(indexes), ptr(cam_pos), v, channels, degree, ptr(features),
           stream)
    else:
      assert feature_c.ndim == 2, f"Features must be (N, C) if use_sh=False, got {feature_c.shape}"
      features = feature_c[indexes]
    F = features.shape[1]

    # accumulated outputs (zeroed) and the raster digest
    image = torch.empty((h, w, F), dtype=dtype, device=device)
    alpha = torch.empty((h, w), dtype=dtype, device=device)
    heuristic = torch.empty((v, 2) if config.compute_point_heuristic else (0, 2), dtype=dtype, device=device)
    visibility = torch.empty((v,) if config.compute_visibility else (0,), dtype=dtype, device=device)
    median = torch.empty((0,), dtype=dtype, device=device)
    digest = torch.empty((0, 16), dtype=torch.float32, device=device)
    fused_median = render_median_depth and fused_median_supported(config, F, dtype)
    use_digest = tuned_supported(config, F, dtype) and (fused_median or not render_median_depth)
    heuristic.zero_()
    visibility.zero_()
    if use_digest:   # raster records, written once and gathered by the forward and the backward kernel
      digest = torch.empty((v, 16), dtype=torch.float32, device=device)
      if v > 0:
        call("gs_raster_digest_f32", ptr(g2d), ptr(features), ptr(depths) if fused_median else None, v, F,
             _lib.raster_config_c(config), ptr(digest), stream)

    # ---- tile mapper (fp32 only, like the reference): two-level ordering, one host read (K) ----
    g32 = g2d if dtype == torch.float32 else g2d.float()
    d32 = (ndc if dtype == torch.float32 else ndc.float()).view(-1)
    binned = bin_and_sort_binned(g32, d32, (w, h), config, use_depth16) if ORDERING == "binned" else None
    if binned is not None:
      overlap_to_point, tile_ranges = binned
    else:
      overlap_to_point, tile_ranges, _, _, _ = bin_and_sort(g32, d32, (w, h), config, use_depth16)
    k = overlap_to_point.shape[0]
    ranges = tile_ranges.view(-1, 2)

    # ---- rasteriser (+ fused median depth) ----
    vis_ptr = ptr(visibility) if config.compute_visibility else None
    cfg = _lib.raster_config_c(config)
    packed = None
    if use_digest:
      if fused_median:
        median = torch.empty((h, w), dtype=dtype, device=device)
      if config.use_alpha_blending:   # per-overlap records in sorted order, then the bulk-copy staged kernel
        packed = pack_records(digest, ranges, overlap_to_point, (w, h), F)
        call("gs_raster_fwd_packed_f32", ptr(packed[0]), ptr(ranges), ptr(overlap_to_point), v, k, w, h, F, cfg,
             float(config.median_threshold), ptr(image), ptr(alpha), vis_ptr, ptr(median) if fused_median else None,
             stream)
      else:
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
        median_alpha = torch.empty((h, w), dtype=dtype, device=device)   # written, not used; named so it outlives the launch
        call(f"gs_raster_fwd_{sfx}", ptr(g2d), ptr(depths), ptr(ranges), ptr(overlap_to_point), v, k, w, h, 1, dcfg,
             ptr(median3), ptr(median_alpha), None, stream)
        median = median3.squeeze(-1)

    ctx.save_for_backward(*tensors, feature_c, indexes, g2d, features, image, overlap_to_point, ranges,
                          cam_pos if cam_pos is not None else torch.empty(0, device=device), digest)
    ctx.packed = packed
    ctx.meta = (config, (w, h), blur, margin, bool(use_sh), heuristic)
    ctx.sh_exchange = sh_exchange
    ctx.set_materialize_grads(False)
    ctx.mark_non_differentiable(alpha, indexes, visibility, heuristic, median, overlap_to_point, tile_ranges)
    return image, alpha, g2d, depths, indexes, features, visibility, heuristic, median, overlap_to_point, tile_ranges

  @staticmethod
  def _forward_fused_host(ctx, position, log_scaling, rotation, alpha_logit, feature, T_camera_world, projection,
                          camera, config, use_sh, use_depth16, render_median_depth, sh_exchange):
    """Same stages through gs_render_stage_a/b_f32: two C calls instead of ~20, buffers still allocated here."""
    device = position.device
    f32, i32 = torch.float32, torch.int32
    ptr = _lib.ptr
    stream = _lib.stream_ptr(device)
    n = position.shape[0]
    w, h = int(camera.image_size[0]), int(camera.image_size[1])
    tensors = [t.detach().contiguous() for t in (position, log_scaling, rotation, alpha_logit, T_camera_world, projection)]
    feature_c = feature.detach().contiguous()
    F = feature_c.shape[1]
    degree = check_sh_degree(feature_c) if use_sh else 0
    ts = config.tile_size
    tile_shape = ((h + ts - 1) // ts, (w + ts - 1) // ts)
    assert tile_shape[0] * tile_shape[1] < 65535, \
        f"tile dimensions {tile_shape} for image size {(w, h)} exceed maximum tile count (16 bit id), try increasing tile_size"

    def empty(shape, dtype=f32):
      return torch.empty(shape, dtype=dtype, device=device)

    # capacity-n buffers (V is only known inside stage A); narrowed to V rows below
    g2d_n, depths_n, ndc_n, idx_n = empty((n, 7)), empty((n, 1)), empty((n, 1)), empty((n,), torch.int64)
    feat_n, digest_n = empty((n, F)), empty((n, 16))
    vis_n = empty((n,)) if config.compute_visibility else None
    heur_n = empty((n, 2)) if config.compute_point_heuristic else None
    cam_pos = empty((3,))
    order, counts, cum = empty((n,), i32), empty((n,), i32), empty((n + 1,), i32)
    hits = empty((n, 2), torch.int64)   # per-Gaussian hit records shared by tile count and emit
    ws_bytes = _workspace_sizes(n)
    ws = [_lib.workspace(b, device) for b in ws_bytes]
    image, alpha = empty((h, w, F)), empty((h, w))
    median = empty((h, w)) if render_median_depth else empty((0,))
    tile_ranges = empty((*tile_shape, 2), i32)
    num_tiles = tile_shape[0] * tile_shape[1]
    tile_counts, tile_cursor, tile_totals = empty((num_tiles,), i32), empty((num_tiles,), i32), empty((2,), i32)
    ev_fwd = _next_event_pair("fwd")

    args = _lib.RenderArgsC(
        ptr(tensors[0]), ptr(tensors[1]), ptr(tensors[2]), ptr(tensors[3]), ptr(feature_c), ptr(tensors[4]), ptr(tensors[5]),
        n, w, h, float(camera.near_plane), float(camera.far_plane), float(config.blur_cov), float(config.clamp_margin),
        float(config.median_threshold), int(use_sh), degree, F, int(use_depth16), int(render_median_depth),
        int(ORDERING == "binned"),
        _lib.raster_config_c(config),
        ptr(g2d_n), ptr(depths_n), ptr(ndc_n), ptr(idx_n), ptr(feat_n), ptr(digest_n),
        ptr(vis_n), ptr(heur_n), ptr(cam_pos), ptr(order), ptr(counts), ptr(cum),
        ws[0].data_ptr(), ws[0].numel(), ws[1].data_ptr(), ws[1].numel(), ws[2].data_ptr(), ws[2].numel(),
        ptr(image), ptr(alpha), ptr(median) if render_median_depth else None, ptr(tile_ranges),
        _event_handle(ev_fwd[0] if ev_fwd else None), _event_handle(ev_fwd[1] if ev_fwd else None),
        ptr(tile_counts), ptr(tile_cursor), ptr(tile_totals), None, None, ptr(hits), 0, 0)
    tile_shard = sh_exchange if getattr(sh_exchange, "kind", None) == "tile" else None
    if tile_shard is not None:   # this rank bins, sorts, packs and rasterises its own contiguous tile-id range only
      assert ORDERING != "binned", "tile sharding uses the two-level ordering"
      args.tile_lo, args.tile_hi = tile_shard.tile_range(num_tiles)
    rec_cols = 12 if F <= 3 else 16   # floats per packed raster record (gs_raster_pack_bytes)
    # K-sized buffers: sized from the previous frame on this device (+25 %), so that the driver can go from the
    # host read of K straight into key emission; if K outgrew them, allocate exactly and run stage B from here
    v_out, k_out, max_out, done = _lib.c_int64(), _lib.c_int64(), _lib.c_int64(), _lib.c_int32()
    nbytes = _lib.c_size_t()
    dev_key = device.index if device.index is not None else torch.cuda.current_device()
    cap = _k_capacity.get(dev_key, 0)
    tiles = o2p = ws_sort = records = flush_records = None
    if cap > 0:
      tiles, o2p = empty((2, cap), i32), empty((2, cap), i32)
      records, flush_records = empty((cap, rec_cols)), empty((cap, 4))
      args.records, args.flush_records = ptr(records), ptr(flush_records)
      _lib.call("gs_sort_pairs_workspace_bytes", cap, 4, nbytes)
      ws_sort = _lib.workspace(nbytes.value, device)
    _lib.call("gs_render_forward_f32", args, cap, ptr(tiles), ptr(o2p), ws_sort.data_ptr() if cap > 0 else None,
              ws_sort.numel() if cap > 0 else 0, v_out, k_out, max_out, done, stream)
    v, k = int(v_out.value), int(k_out.value)
    if done.value and _lib.profiler is not None:
      _lib.profiler.launches += _lib.OWN_KERNELS["gs_render_stage_b_f32"]   # stage B ran inside the forward driver
    if not done.value:
      cap = k
      tiles, o2p = empty((2, k), i32), empty((2, k), i32)
      records, flush_records = empty((k, rec_cols)), empty((k, 4))
      args.records, args.flush_records = ptr(records), ptr(flush_records)
      _lib.call("gs_sort_pairs_workspace_bytes", k, 4, nbytes)
      ws_sort = _lib.workspace(nbytes.value, device)
      _lib.call("gs_render_stage_b_f32", args, v, k, int(max_out.value), k, ptr(tiles), ptr(o2p), ws_sort.data_ptr(),
                ws_sort.numel(), stream)
    if o2p is None:   # first frame on this device and nothing to rasterise (K = 0): stage B ran with empty buffers
      tiles, o2p = empty((2, 0), i32), empty((2, 0), i32)
      records, flush_records = empty((0, rec_cols)), empty((0, 4))
    _k_capacity[dev_key] = max(int(k * 1.25), 1024)

    g2d, depths, indexes, features, digest = g2d_n[:v], depths_n[:v], idx_n[:v], feat_n[:v], digest_n[:v]
    visibility = vis_n[:v] if vis_n is not None else empty((0,))
    heuristic = heur_n[:v] if heur_n is not None else empty((0, 2))
    overlap_to_point, ranges = o2p[1, :k], tile_ranges.view(-1, 2)
    ctx.save_for_backward(*tensors, feature_c, indexes, g2d, features, image, overlap_to_point, ranges, cam_pos, digest)
    ctx.packed = (records[:k], flush_records[:k])   # per-overlap raster records of this frame: the backward sweeps them again
    ctx.meta = (config, (w, h), float(config.blur_cov), float(config.clamp_margin), bool(use_sh), heuristic)
    ctx.sh_exchange = sh_exchange
    ctx.set_materialize_grads(False)
    ctx.mark_non_differentiable(alpha, indexes, visibility, heuristic, median, overlap_to_point, tile_ranges)
    return image, alpha, g2d, depths, indexes, features, visibility, heuristic, median, overlap_to_point, tile_ranges

  @staticmethod
  def _backward_fused_host(ctx, d_image, d_g2d, d_depths, d_features, tile_shard=None):
    """The backward through gs_render_backward_f32.  tile_shard (tile-sharded multi-GPU run): the raster backward
    leaves this rank's partial sums of the packed-2D / feature gradients (its own tiles only) in ONE flat buffer,
    which is summed over ranks by a single NCCL all-reduce before the replicated SH / projection backward."""
    (position, log_scaling, rotation, alpha_logit, T_camera_world, projection, feature, indexes, g2d, features, image,
     overlap_to_point, ranges, cam_pos, digest) = ctx.saved_tensors
    config, (w, h), blur, margin, use_sh, heuristic = ctx.meta
    device = position.device
    ptr = _lib.ptr
    n, v, k, F = position.shape[0], g2d.shape[0], overlap_to_point.shape[0], features.shape[1]
    need = ctx.needs_input_grad
    grads = [torch.empty_like(t) if need[i] else None
             for t, i in ((position, 0), (log_scaling, 1), (rotation, 2), (alpha_logit, 3), (T_camera_world, 5), (projection, 6))]
    d_feature = torch.empty_like(feature) if need[4] else None
    if tile_shard is not None:
      flat = torch.empty((v * (7 + F),), dtype=torch.float32, device=device)
      grad_g, grad_f = flat[:7 * v].view(v, 7), flat[7 * v:].view(v, F)
      if d_g2d is not None:        # incoming gradients enter the sum once (they are replicated over ranks)
        grad_g.copy_(d_g2d if tile_shard.rank == 0 else torch.zeros_like(d_g2d))
      if d_features is not None:
        grad_f.copy_(d_features if tile_shard.rank == 0 else torch.zeros_like(d_features))
    else:
      grad_g = d_g2d.clone() if d_g2d is not None else torch.empty_like(g2d)
      grad_f = (d_features.clone() if d_features is not None else torch.empty_like(features)) if need[4] else None
    # the SH view directions depend on the camera centre inverse(T)[:3,3]: when the pose is trained, its gradient
    # through them is chained into d_T_camera_world below (reference: autograd through camera_params.camera_position)
    d_cam = torch.empty((3,), dtype=torch.float32, device=device) if (use_sh and need[4] and need[5]) else None
    ev_bwd = _next_event_pair("bwd")
    # dL/dimage goes to the kernel with whatever strides autograd gave it (an expanded scalar after image.sum(), a
    # permuted CHW tensor, ...): no .contiguous() copy of a full image
    strided = d_image is not None and not d_image.is_contiguous()
    d_image_strides = (_lib.c_int64 * 3)(*(d_image.stride() if strided else (0, 0, 0)))
    d_depths_c = d_depths.contiguous() if d_depths is not None else None   # named: outlives the launches below
    args = _lib.RenderBwdArgsC(
        ptr(position), ptr(log_scaling), ptr(rotation), ptr(alpha_logit), ptr(feature), ptr(T_camera_world), ptr(projection),
        n, v, k, w, h, blur, margin, int(use_sh), check_sh_degree(feature) if use_sh else 0, F, int(strided),
        _lib.raster_config_c(config),
        ptr(indexes), ptr(features), ptr(image), ptr(cam_pos), ptr(digest), ptr(overlap_to_point), ptr(ranges),
        ptr(ctx.packed[0]), ptr(ctx.packed[1]),
        (d_image.data_ptr() if strided else ptr(d_image)) if d_image is not None else None,
        ptr(d_depths_c),
        ptr(grad_g), ptr(grad_f), int(d_g2d is not None), int(d_features is not None and grad_f is not None),
        ptr(heuristic) if config.compute_point_heuristic else None,
        *[ptr(g) for g in grads], ptr(d_feature),
        _event_handle(ev_bwd[0] if ev_bwd else None), _event_handle(ev_bwd[1] if ev_bwd else None), d_image_strides,
        ptr(d_cam), 0)
    if tile_shard is not None and tile_shard.world > 1:
      args.phases = _lib.GS_BWD_RASTER
      _lib.call("gs_render_backward_f32", args, _lib.stream_ptr(device))
      tile_shard.reduce(flat)      # THE collective of the tile-sharded path: 4 (7 + F) bytes per visible Gaussian
      args.phases = _lib.GS_BWD_FEATURE | _lib.GS_BWD_PROJECT
      if _lib.profiler is not None:   # the two phase calls together launch the driver's 3 kernels once
        _lib.profiler.launches -= _lib.OWN_KERNELS["gs_render_backward_f32"]
    _lib.call("gs_render_backward_f32", args, _lib.stream_ptr(device))
    if d_cam is not None:
      grads[4] += camera_position_vjp(T_camera_world, cam_pos, d_cam)
    return (grads[0], grads[1], grads[2], grads[3], d_feature, grads[4], grads[5], None, None, None, None, None, None)

  @staticmethod
  def _backward_fused_view_parallel(ctx, d_image, d_g2d, d_depths, d_features):
    """View-parallel backward through gs_render_backward_f32 in two phases with the exchange between them:
       raster backward (C) -> pack + all-gather of the SH factors (NCCL, async) -> projection backward (C) into ONE
       flat geometry buffer -> all-reduce of that buffer (NCCL, async) -> rebuild the SH gradient from the gathered
       factors (overlaps the all-reduce) -> wait."""
    import torch.distributed as dist
    (position, log_scaling, rotation, alpha_logit, T_camera_world, projection, feature, indexes, g2d, features, image,
     overlap_to_point, ranges, cam_pos, digest) = ctx.saved_tensors
    config, (w, h), blur, margin, use_sh, heuristic = ctx.meta
    exchange = ctx.sh_exchange
    device = position.device
    ptr = _lib.ptr
    n, v, k, F = position.shape[0], g2d.shape[0], overlap_to_point.shape[0], features.shape[1]
    need = ctx.needs_input_grad
    # position | log_scaling | rotation | alpha_logit as consecutive blocks of one buffer: one collective, no
    # concatenation and no copy back
    # with peer memory the flat buffer is a persistent symmetric allocation, all-reduced in place by our own kernel
    geom_state = exchange.geometry_state(n, device) if exchange.reduce_geometry else None
    flat = geom_state["buf"] if geom_state is not None else torch.empty((11 * n,), dtype=torch.float32, device=device)
    geom = [flat[0:3 * n].view(n, 3), flat[3 * n:6 * n].view(n, 3), flat[6 * n:10 * n].view(n, 4), flat[10 * n:11 * n].view(n, 1)]
    d_T = torch.empty_like(T_camera_world) if need[5] else None
    d_proj = torch.empty_like(projection) if need[6] else None
    grad_g = d_g2d.clone() if d_g2d is not None else torch.empty_like(g2d)
    grad_f = d_features.clone() if d_features is not None else torch.empty_like(features)
    ev_bwd = _next_event_pair("bwd")
    strided = d_image is not None and not d_image.is_contiguous()
    d_image_strides = (_lib.c_int64 * 3)(*(d_image.stride() if strided else (0, 0, 0)))
    d_depths_c = d_depths.contiguous() if d_depths is not None else None   # named: outlives the launches below
    args = _lib.RenderBwdArgsC(
        ptr(position), ptr(log_scaling), ptr(rotation), ptr(alpha_logit), ptr(feature), ptr(T_camera_world), ptr(projection),
        n, v, k, w, h, blur, margin, int(use_sh), check_sh_degree(feature), F, int(strided),
        _lib.raster_config_c(config),
        ptr(indexes), ptr(features), ptr(image), ptr(cam_pos), ptr(digest), ptr(overlap_to_point), ptr(ranges),
        ptr(ctx.packed[0]), ptr(ctx.packed[1]),
        (d_image.data_ptr() if strided else ptr(d_image)) if d_image is not None else None,
        ptr(d_depths_c),
        ptr(grad_g), ptr(grad_f), int(d_g2d is not None), int(d_features is not None),
        ptr(heuristic) if config.compute_point_heuristic else None,
        *[ptr(g) for g in geom], ptr(d_T), ptr(d_proj), None,
        _event_handle(ev_bwd[0] if ev_bwd else None), _event_handle(ev_bwd[1] if ev_bwd else None), d_image_strides,
        None, _lib.GS_BWD_RASTER)
    stream = _lib.stream_ptr(device)
    _lib.call("gs_render_backward_f32", args, stream)
    pending = exchange.start(feature, indexes, features, grad_f, cam_pos)
    args.phases = _lib.GS_BWD_PROJECT
    _lib.call("gs_render_backward_f32", args, stream)
    if _lib.profiler is not None:   # two phase calls, each counted as the whole driver (3 kernels): raster + projection = 2
      _lib.profiler.launches -= 2 * _lib.OWN_KERNELS["gs_render_backward_f32"] - 2
    if geom_state is not None and pending[0] == "peer":
      d_feature, flat = exchange.finish_with_geometry(pending, geom_state, feature, position, check_sh_degree(feature))
      geom = [flat[0:3 * n].view(n, 3), flat[3 * n:6 * n].view(n, 3), flat[6 * n:10 * n].view(n, 4), flat[10 * n:11 * n].view(n, 1)]
      grads = [g if need[i] else None for i, g in enumerate(geom)]
      return (grads[0], grads[1], grads[2], grads[3], d_feature, d_T, d_proj, None, None, None, None, None, None)
    # NCCL path.  torch's own symmetric-memory all-reduce is no alternative: torch's symmetric-memory multimem all-reduce of this 44 MB buffer was measured
    # at 0.40 ms on 2 GPUs against 0.11-0.17 ms for the NCCL ring (profiles/r02/r02o_timeline_n2.txt)
    reduce = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=exchange.group, async_op=True) if exchange.reduce_geometry else None
    d_feature = exchange.finish(pending, feature, position, check_sh_degree(feature))
    if reduce is not None:
      reduce.wait()
    grads = [g if need[i] else None for i, g in enumerate(geom)]
    return (grads[0], grads[1], grads[2], grads[3], d_feature, d_T, d_proj, None, None, None, None, None, None)

  @staticmethod
  def backward(ctx, d_image, d_alpha, d_g2d, d_depths, d_indexes, d_features, *unused):
    if ctx.fused_host:
      exchange = ctx.sh_exchange
      if getattr(exchange, "kind", None) == "tile":
        return _RenderFunction._backward_fused_host(ctx, d_image, d_g2d, d_depths, d_features, tile_shard=exchange)
      if exchange is None or exchange.world <= 1 or not (ctx.needs_input_grad[4] and ctx.meta[4]):
        return _RenderFunction._backward_fused_host(ctx, d_image, d_g2d, d_depths, d_features)
      if os.environ.get("GS_VIEW_PARALLEL_STAGED", "0") != "1":
        return _RenderFunction._backward_fused_view_parallel(ctx, d_image, d_g2d, d_depths, d_features)
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

    exchange = ctx.sh_exchange if (need[4] and use_sh and dtype == torch.float32) else None
    if exchange is not None and exchange.world <= 1:
      exchange = None

    grads = [torch.empty_like(t) if need[i] else None
             for t, i in ((position, 0), (log_scaling, 1), (rotation, 2), (alpha_logit, 3), (T_camera_world, 5), (projection, 6))]
    sh_direct = need[4] and exchange is None
    all_rows_written = use_sh and v == feature.shape[0]
    d_feature = torch.empty_like(feature) if sh_direct else None
    for t in grads:
      if t is not None:
        t.zero_()
    if sh_direct and not all_rows_written:
      d_feature.zero_()

    # ---- rasteriser backward: gradients of the packed 2D Gaussians and per-point features ----
    grad_g = d_g2d.clone() if d_g2d is not None else torch.zeros_like(g2d)
    grad_f = d_features.clone() if d_features is not None else torch.zeros_like(features)
    if d_image is not None and v > 0:
      d_image_c = d_image.contiguous()   # named: the pointer must not outlive a temporary copy
      out_ptrs = (ptr(grad_g) if need_geom else None, ptr(grad_f) if need[4] else None,
                  ptr(heuristic) if config.compute_point_heuristic else None)
      if getattr(ctx, "packed", None) is not None:
        call("gs_raster_bwd_packed_f32", ptr(ctx.packed[0]), ptr(ctx.packed[1]), ptr(ranges), ptr(overlap_to_point),
             ptr(image), ptr(d_image_c), None, v, overlap_to_point.shape[0], w, h, F,
             _lib.raster_config_c(config), *out_ptrs, stream)
      elif digest.shape[0] == v:
        call("gs_raster_bwd_digest_f32", ptr(digest), ptr(ranges), ptr(overlap_to_point), ptr(image),
             ptr(d_image_c), v, overlap_to_point.shape[0], w, h, F, _lib.raster_config_c(config),
             *out_ptrs, stream)
      else:
        call(f"gs_raster_bwd_{sfx}", ptr(g2d), ptr(features), ptr(ranges), ptr(overlap_to_point), ptr(image),
             ptr(d_image_c), v, overlap_to_point.shape[0], w, h, F, _lib.raster_config_c(config),
             *out_ptrs, stream)

    # ---- features: SH backward; view-parallel runs launch the exchange of the SH-gradient factors instead, so that
    # the all-gather overlaps the projection backward ----
    pending = exchange.start(feature, indexes, features, grad_f, cam_pos) if exchange is not None else None
    d_cam = None
    if sh_direct and v > 0:
      if use_sh:
        d_cam = torch.zeros((3,), dtype=dtype, device=device) if need[5] else None
        call(f"gs_sh_bwd_{sfx}", ptr(feature), ptr(position), ptr(indexes), ptr(cam_pos), ptr(grad_f), ptr(features), v,
             feature.shape[1], check_sh_degree(feature), 1, ptr(d_feature), None, ptr(d_cam), stream)
      else:
        d_feature.index_copy_(0, indexes, grad_f)

    # ---- projection backward ----
    if need_geom and v > 0:
      dd = d_depths.contiguous() if d_depths is not None else torch.zeros((v, 1), dtype=dtype, device=device)
      call(f"gs_project_bwd_{sfx}", ptr(position), ptr(log_scaling), ptr(rotation), ptr(alpha_logit), ptr(T_camera_world),
           ptr(projection), ptr(indexes), v, w, h, blur, margin, ptr(grad_g), ptr(dd), *[ptr(g) for g in grads], stream)

    if d_cam is not None:   # SH view directions -> camera centre -> view matrix (reference: autograd through torch.inverse)
      grads[4] += camera_position_vjp(T_camera_world, cam_pos, d_cam)
    if exchange is not None:
      d_feature = exchange.finish(pending, feature, position, check_sh_degree(feature))
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
