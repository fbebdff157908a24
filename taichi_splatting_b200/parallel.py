"""Multi-GPU sharding of the render path (new design: the reference has no distributed code).

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch; gloo on CPU for the host-logic tests).
The path shards two natural ways, both with exactly ONE collective, in the backward pass:

  view-parallel   the cloud is replicated, rank r renders its own camera(s); no communication in the
                  forward; the backward sums the per-Gaussian parameter gradients over ranks: either one
                  all-reduce of everything (`allreduce_gradients`), or -- `render_view_parallel` -- an
                  all-reduce of the geometry gradients (44 B / Gaussian) plus an all-gather of the rank-1
                  factors of the SH gradient (12 B / Gaussian / view instead of 192 B; `ShGradientExchange`).
  tile-sharded    one view, the tile grid is cut into `world` contiguous tile-id ranges holding equal
                  numbers of (tile, Gaussian) overlaps (`partition_tiles`); every rank projects / bins / sorts
                  the whole (cheap, O(N)) front end, rasterises only its own tiles, and the gradients of the
                  packed 2D Gaussians and features are summed across ranks (`reduce_across_ranks`, an identity
                  in the forward) before the replicated projection / SH backward.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world_info(group=None) -> Tuple[int, int]:
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(group), dist.get_world_size(group)
  return 0, 1


# ------------------------------------------------------------------------------------------ view-parallel
def views_for_rank(num_views: int, rank: int, world: int) -> List[int]:
  """Round-robin assignment of camera views to ranks."""
  return list(range(rank, num_views, world))


def allreduce_gradients(tensors: Sequence[torch.Tensor], group=None, bucket: bool = True) -> None:
  """Sum the `.grad` of every tensor over all ranks, in place: the single exchange of the view-parallel path.

  With `bucket` the gradients travel as one flat buffer (one collective launch; NVSwitch makes the cost
  latency- not link-bound, so fewer, larger messages win); tensors without a gradient contribute zeros."""
  rank, world = world_info(group)
  if world == 1:
    return
  grads = []
  for t in tensors:
    if t.grad is None:
      t.grad = torch.zeros_like(t)
    grads.append(t.grad)
  if not bucket or len(grads) == 1:
    works = [dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=True) for g in grads]
    for w in works:
      w.wait()
    return
  # gradients that already sit back to back in one allocation (the view-parallel backward writes the geometry
  # gradients into one flat buffer): reduce that range in place, no concatenation and no copy back
  adjacent = all(g.is_contiguous() and g.dtype == grads[0].dtype for g in grads) and all(
      b.untyped_storage().data_ptr() == a.untyped_storage().data_ptr()
      and b.data_ptr() == a.data_ptr() + a.numel() * a.element_size() for a, b in zip(grads, grads[1:]))
  if adjacent:
    total = sum(g.numel() for g in grads)
    flat = grads[0].new_empty(0).set_(grads[0].untyped_storage(), grads[0].storage_offset(), (total,), (1,))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return
  flat = torch.cat([g.reshape(-1) for g in grads])
  dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
  offset = 0
  for g in grads:
    n = g.numel()
    g.copy_(flat[offset:offset + n].view_as(g))
    offset += n


# symmetric-memory allocations (gathered factor slots, flat geometry buffer) are collective and cost ~100 ms to set up:
# one per (process group, shape, device) for the life of the process, shared by every ShGradientExchange object
_PEER_STATE = {}


class ShGradientExchange:
  """View-parallel exchange of the spherical-harmonics gradient, exploiting its structure.

  For one view the SH coefficient gradient is rank-1 per Gaussian and channel,
      d_params[i, c, :] = Y(normalize(p_i - camera)) * g[i, c],     g = dL/dcolour (0 where clamped / culled),
  so instead of all-reducing the (N, C, 16) tensor (192 B per Gaussian at degree 3, 81 % of all gradient bytes)
  every rank all-gathers its dense (N, C) factor g and its camera centre (12 B per Gaussian and view) and rebuilds
  sum_w Y_w * g_w locally with one kernel (gs_sh_bwd_views_f32).  The result equals the all-reduced gradient up to
  fp32 summation order.  Used through `render_view_parallel`; the renderer's backward calls `sum_sh_gradient`."""

  def __init__(self, group=None, reduce_geometry=False, peer_memory: Optional[bool] = None):
    self.group = group
    self.rank, self.world = world_info(group)
    # reduce_geometry: the backward also all-reduces the geometry gradients (position, log_scaling, rotation,
    # alpha_logit), written by the projection backward into ONE flat buffer, while the SH gradient is rebuilt -- the
    # caller then needs no collective of its own after loss.backward()
    self.reduce_geometry = reduce_geometry
    # peer_memory: the all-gather of the factors is FUSED into the kernel that packs them -- it stores straight into
    # every rank's gathered buffer over NVLink (symmetric memory) -- instead of pack + NCCL all-gather; the geometry
    # buffer is then all-reduced in place by gs_allreduce_peers_f32.  None: only when GS_PEER_EXCHANGE=1.
    self.peer_memory = peer_memory
    self._peer = _PEER_STATE.setdefault(id(group) if group is not None else 0, {})   # shared across frames

  def _peer_state(self, n, channels, device):
    """Two gathered slots (alternating per frame, so that a fast rank writing frame i + 1 never touches what a slow
    rank still reads for frame i; one barrier per frame orders the rest) in symmetric memory, mapped on every rank."""
    import os
    key = (n, channels, device.index)
    state = None
    # default OFF: at 8 ranks the NCCL exchange measured 2.53 ms per step against 2.58-2.64 with peer memory (equal at
    # 2 and 4 ranks; profiles/r02/r02q_*, r02s_*) -- opt in with peer_memory=True or GS_PEER_EXCHANGE=1
    want = self.peer_memory if self.peer_memory is not None else os.environ.get("GS_PEER_EXCHANGE", "0") == "1"
    if not (want and self.world > 1 and device.type == "cuda" and dist.get_backend(self.group) == "nccl"):
      return None
    if key in self._peer:
      return self._peer[key]
    try:
      import torch.distributed._symmetric_memory as symm_mem
      stride = n * channels + 3
      buf = symm_mem.empty((2, self.world, stride), dtype=torch.float32, device=device)
      handle = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
      bases = [int(p) for p in handle.buffer_ptrs]
      assert len(bases) == self.world and all(bases)
      state = dict(buf=buf, handle=handle, bases=bases, stride=stride, frame=0, stream=torch.cuda.Stream(device=device))
    except Exception as e:   # no peer mapping on this system: the NCCL path is always there
      if self.peer_memory:
        raise
      state = None
      self._peer_error = repr(e)
    self._peer[key] = state
    return state

  def start(self, sh_params, indexes, colours, d_colours, camera_pos):
    """Make this rank's factors available on every rank (asynchronously: the caller overlaps the projection backward).
    Peer-memory mode: ONE kernel packs them (clamp mask, scatter to dense rows, camera centre) and stores them into
    every rank's gathered buffer (gs_sh_pack_factors_peers_f32).  Otherwise: gs_sh_pack_factors_f32 + NCCL all-gather."""
    from . import _lib
    n, channels = sh_params.shape[0], sh_params.shape[1]
    device = sh_params.device
    peer = self._peer_state(n, channels, device)
    if peer is not None:
      slot = peer["frame"] % 2
      peer["frame"] += 1
      stride = peer["stride"]
      bases = (_lib.ctypes.c_uint64 * self.world)(*peer["bases"])
      # on its own stream: the stores to the peers (NVLink egress, 0.12 ms at 8 ranks) run beside the projection backward
      main, side = torch.cuda.current_stream(device), peer["stream"]
      side.wait_stream(main)
      with torch.cuda.stream(side):
        _lib.call("gs_sh_pack_factors_peers_f32", _lib.ptr(colours), _lib.ptr(d_colours), _lib.ptr(indexes),
                  _lib.ptr(camera_pos), indexes.shape[0], channels, n, bases, self.world,
                  (slot * self.world + self.rank) * stride, side.cuda_stream, on=side)
        peer["pack_done"] = torch.cuda.Event()
        peer["pack_done"].record(side)
      for t in (colours, d_colours, indexes, camera_pos):
        t.record_stream(side)
      return ("peer", peer, slot, stride)
    stride = n * channels + 3
    local = torch.empty((stride,), dtype=torch.float32, device=device)
    _lib.call("gs_sh_pack_factors_f32", _lib.ptr(colours), _lib.ptr(d_colours), _lib.ptr(indexes), _lib.ptr(camera_pos),
              indexes.shape[0], channels, n, _lib.ptr(local), _lib.stream_ptr(device))
    gathered = torch.empty((self.world, stride), dtype=torch.float32, device=device)
    work = dist.all_gather_into_tensor(gathered, local, group=self.group, async_op=True)
    return (work, gathered, local, stride)

  # ---- geometry gradients in peer memory: in-place all-reduce by gs_allreduce_peers_f32 ---------------------------
  def geometry_state(self, n: int, device):
    """Persistent symmetric (11 n, padded to a multiple of 4) buffer for position | log_scaling | rotation | alpha_logit
    gradients, or None when peer memory is not in use (first frame included: the factor buffers are set up first)."""
    import os
    want = self.peer_memory if self.peer_memory is not None else os.environ.get("GS_PEER_EXCHANGE", "0") == "1"
    if not want or not any(st is not None and k[0] != "geom" and k[-1] == device.index for k, st in self._peer.items()):
      return None
    key = ("geom", n, device.index)
    if key not in self._peer:
      import torch.distributed._symmetric_memory as symm_mem
      count = (11 * n + 3) // 4 * 4
      buf = symm_mem.empty((count,), dtype=torch.float32, device=device)
      buf.zero_()
      handle = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
      self._peer[key] = dict(buf=buf, handle=handle, bases=[int(p) for p in handle.buffer_ptrs], count=count,
                             stream=torch.cuda.Stream(device=device))
    return self._peer[key]

  def finish_with_geometry(self, pending, geom_state, sh_params, positions, degree):
    """Peer-memory tail of the view-parallel backward (after the projection backward has been enqueued on the current
    stream): ONE barrier says that every rank's factors are stored everywhere and every rank's geometry gradients are
    written; then the SH rebuild (current stream) runs beside the in-place peer all-reduce of the geometry buffer
    (side stream, closed by a second barrier).  Returns (d_params, reduced geometry buffer copy)."""
    from . import _lib
    _, peer, slot, _stride = pending
    device = sh_params.device
    main, side = torch.cuda.current_stream(device), geom_state["stream"]
    if peer.get("pack_done") is not None:
      main.wait_event(peer["pack_done"])
    peer["handle"].barrier(channel=slot)
    side.wait_stream(main)
    with torch.cuda.stream(side):
      bases = (_lib.ctypes.c_uint64 * self.world)(*geom_state["bases"])
      _lib.call("gs_allreduce_peers_f32", bases, self.world, self.rank, geom_state["count"], side.cuda_stream, on=side)
      geom_state["handle"].barrier(channel=0)
      reduced = geom_state["buf"].clone()      # autograd gets its own copy; the symmetric buffer serves the next frame
    n, channels = sh_params.shape[0], sh_params.shape[1]
    gathered = peer["buf"][slot]
    cams = gathered[:, n * channels:].contiguous()
    d_params = torch.empty_like(sh_params)
    _lib.call("gs_sh_bwd_views_f32", _lib.ptr(positions), _lib.ptr(cams), _lib.ptr(gathered), n, self.world, channels,
              peer["stride"], degree, _lib.ptr(d_params), _lib.stream_ptr(device))
    main.wait_stream(side)
    reduced.record_stream(main)
    return d_params, reduced

  def finish(self, pending, sh_params, positions, degree):
    """Wait for the gathered factors and rebuild sum_w Y_w * g_w -> d_params (N, C, D)."""
    from . import _lib
    if pending[0] == "peer":
      _, peer, slot, stride = pending
      if peer.get("pack_done") is not None:
        torch.cuda.current_stream(sh_params.device).wait_event(peer["pack_done"])
      # every rank's stores into every gathered buffer are complete and visible once all ranks have passed this
      # barrier (signal pads of the symmetric allocation, enqueued on the current stream: no host involvement)
      peer["handle"].barrier(channel=slot)
      gathered = peer["buf"][slot]
    else:
      work, gathered, _local, stride = pending
      work.wait()
    n, channels = sh_params.shape[0], sh_params.shape[1]
    cams = gathered[:, n * channels:].contiguous()
    d_params = torch.empty_like(sh_params)
    _lib.call("gs_sh_bwd_views_f32", _lib.ptr(positions), _lib.ptr(cams), _lib.ptr(gathered), n, self.world, channels,
              stride, degree, _lib.ptr(d_params), _lib.stream_ptr(sh_params.device))
    return d_params

  def sum_sh_gradient(self, sh_params, positions, indexes, colours, d_colours, camera_pos, degree):
    return self.finish(self.start(sh_params, indexes, colours, d_colours, camera_pos), sh_params, positions, degree)


def render_view_parallel(gaussians, camera_params, config, use_sh: bool = False, use_depth16: bool = False,
                         render_median_depth: bool = False, group=None, reduce_in_backward: bool = False):
  """render_gaussians for this rank's view of a replicated cloud.

  reduce_in_backward=False: after `loss.backward()` call `finish_view_parallel_backward(gaussians, use_sh, group)`:
  with SH the feature gradient has already been summed over ranks inside the backward (ShGradientExchange);
  everything else is all-reduced there.
  reduce_in_backward=True (SH, fp32): the backward itself leaves every per-Gaussian gradient summed over ranks -- the
  all-gather of the SH factors runs beside the projection backward, the all-reduce of the flat geometry-gradient
  buffer beside the kernel that rebuilds the SH gradient -- and nothing remains to be called afterwards (gradients of
  the camera, which differs per rank, stay local)."""
  from .renderer import _RenderFunction, _wrap_rendering
  exchange = ShGradientExchange(group, reduce_geometry=reduce_in_backward) if use_sh else None
  assert exchange is not None or not reduce_in_backward, "reduce_in_backward needs use_sh=True"
  outs = _RenderFunction.apply(*gaussians.shape_tensors(), gaussians.feature, camera_params.T_camera_world,
                               camera_params.projection, camera_params, config, use_sh, use_depth16,
                               render_median_depth, exchange)
  return _wrap_rendering(outs, camera_params, config, render_median_depth)


def finish_view_parallel_backward(gaussians, use_sh: bool, group=None) -> None:
  """The single NCCL all-reduce of the view-parallel path: geometry gradients (44 B per Gaussian), plus the plain
  feature gradient when SH is not used."""
  tensors = [gaussians.position, gaussians.log_scaling, gaussians.rotation, gaussians.alpha_logit]
  if not use_sh or gaussians.feature.dtype != torch.float32:
    tensors.append(gaussians.feature)
  allreduce_gradients(tensors, group=group, bucket=True)   # one 44 B/Gaussian message instead of four


# ------------------------------------------------------------------------------------------ tile-sharded
def partition_tiles(tile_ranges: torch.Tensor, world: int) -> torch.Tensor:
  """Cut the tile grid into `world` contiguous tile-id ranges with (nearly) equal overlap counts.

  tile_ranges: (T,2) or (TH,TW,2) int32 start/end offsets into the sorted overlap list.
  Returns a (world+1,) int64 CPU tensor of tile-id boundaries: rank r owns tiles [b[r], b[r+1])."""
  r = tile_ranges.reshape(-1, 2).to(torch.int64)
  counts = (r[:, 1] - r[:, 0]).clamp_min(0)
  T = counts.shape[0]
  csum = torch.cumsum(counts, 0)
  total = int(csum[-1].item()) if T > 0 else 0
  bounds = [0]
  for k in range(1, world):
    if total == 0:
      b = (T * k) // world
    else:
      target = (total * k + world - 1) // world
      b = int(torch.searchsorted(csum, torch.tensor(target, device=csum.device, dtype=csum.dtype)).item()) + 1
    bounds.append(min(max(b, bounds[-1]), T))
  bounds.append(T)
  return torch.tensor(bounds, dtype=torch.int64)


def mask_tile_ranges(tile_ranges: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
  """Copy of `tile_ranges` with every tile outside [lo, hi) emptied, so a rank's rasteriser CTAs for foreign
  tiles retire immediately (their pixels stay zero and receive no gradient)."""
  flat = tile_ranges.reshape(-1, 2)
  out = torch.zeros_like(flat)
  out[lo:hi] = flat[lo:hi]
  return out.view_as(tile_ranges)


class TileShard:
  """This rank's share of a tile-sharded single view (`render_tile_sharded`).

  The tile grid is cut into `world` contiguous tile-id ranges.  Every rank runs the cheap O(N) front end on the whole
  cloud (projection, SH, depth order) but counts / emits / sorts / packs / rasterises ONLY the overlaps of its own
  tiles (the tile mapper filters by tile id), so the K-sized work is divided by `world`.  The backward sums the
  per-Gaussian gradients of the packed 2D Gaussians and features over ranks with ONE NCCL all-reduce (40 B per visible
  Gaussian at 3 features) and then runs the SH / projection backward replicated, so every rank ends with the full
  parameter gradients.

  Boundaries start as equal tile counts; `rebalance(tile_ranges)` (call it after a frame: one small all-reduce and a
  host read) moves them so that every rank holds the same number of overlaps, using that frame's per-tile counts --
  consecutive frames of a training run see almost the same distribution."""
  kind = "tile"

  def __init__(self, group=None):
    self.group = group
    self.rank, self.world = world_info(group)
    self.bounds = None        # (world + 1,) int64 CPU tensor of tile-id boundaries
    self.num_tiles = None
    self.last_tile_ranges = None   # this rank's (TH,TW,2) tile ranges of the last frame rendered through this shard
    self.last_k = 0                # ... and its number of overlaps

  def tile_range(self, num_tiles: int) -> Tuple[int, int]:
    if self.bounds is None or self.num_tiles != num_tiles:
      self.bounds = torch.tensor([(num_tiles * r) // self.world for r in range(self.world + 1)], dtype=torch.int64)
      self.num_tiles = num_tiles
    return int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])

  def rebalance(self, tile_ranges: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tile_ranges: this rank's (TH,TW,2) ranges of a frame (foreign tiles are (0,0)); default: the last frame rendered
    through this shard.  Returns the new bounds."""
    tile_ranges = tile_ranges if tile_ranges is not None else self.last_tile_ranges
    assert tile_ranges is not None, "rebalance: render a frame through this shard first"
    r = tile_ranges.reshape(-1, 2)
    counts = (r[:, 1] - r[:, 0]).clamp_min(0).to(torch.int64)
    if self.world > 1:
      dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.group)   # tile ownership is disjoint: the sum is the union
    csum = torch.cumsum(counts, 0)
    self.bounds = partition_tiles(torch.stack([csum - counts, csum], dim=1), self.world)
    self.num_tiles = counts.shape[0]
    return self.bounds

  def reduce(self, flat: torch.Tensor) -> None:
    if self.world > 1:
      dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)


class _ReduceGradAcrossRanks(torch.autograd.Function):
  """Identity in the forward; all-reduce(sum) of the incoming gradients in the backward."""

  @staticmethod
  def forward(ctx, group, *tensors):
    ctx.group = group
    return tuple(t.view_as(t) for t in tensors)

  @staticmethod
  def backward(ctx, *grads):
    _, world = world_info(ctx.group)
    if world > 1:
      grads = [g.contiguous() for g in grads]
      works = [dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group, async_op=True) for g in grads]
      for w in works:
        w.wait()
    return (None, *grads)


def reduce_across_ranks(*tensors: torch.Tensor, group=None):
  """Mark tensors whose gradients must be summed over ranks (tile-sharded path: packed 2D Gaussians, features)."""
  return _ReduceGradAcrossRanks.apply(group, *tensors)


def render_tile_sharded(gaussians, camera_params, config, use_sh: bool = False, group=None, shard: Optional[TileShard] = None,
                        use_depth16: bool = False, render_median_depth: bool = False):
  """Single view split over ranks by contiguous tile-id ranges.  Returns (Rendering, (tile_lo, tile_hi)); the
  rendering's image holds this rank's tiles only (zeros elsewhere), so a per-pixel loss can be evaluated locally and
  summed; `points.visibility` / heuristics are this rank's partial sums (sum over ranks = the single-GPU values).

  Pass the same `shard` object every frame (and call `shard.rebalance(...)` now and then) to keep the overlap counts
  of the ranks balanced; without it the tile grid is cut into equal tile counts.  fp32 / tile 16 / alpha blending run
  through the whole-frame drivers (this rank bins, sorts, packs and rasterises only its own tiles); other
  configurations fall back to the operator chain with masked tile ranges."""
  from .renderer import _RenderFunction, _wrap_rendering
  from .rasterizer.function import tuned_supported
  shard = shard if shard is not None else TileShard(group)
  feature = gaussians.feature
  channels = feature.shape[1] if feature.ndim >= 2 else 0
  if (feature.dtype == torch.float32 and config.use_alpha_blending and tuned_supported(config, channels, feature.dtype)
      and feature.ndim == (3 if use_sh else 2)):
    outs = _RenderFunction.apply(*gaussians.shape_tensors(), feature, camera_params.T_camera_world,
                                 camera_params.projection, camera_params, config, use_sh, use_depth16,
                                 render_median_depth, shard)
    ts = config.tile_size
    w, h = camera_params.image_size
    num_tiles = ((w + ts - 1) // ts) * ((h + ts - 1) // ts)
    shard.last_tile_ranges, shard.last_k = outs[-1].detach(), int(outs[-2].shape[0])
    return _wrap_rendering(outs, camera_params, config, render_median_depth), shard.tile_range(num_tiles)
  return _render_tile_sharded_operators(gaussians, camera_params, config, use_sh, shard)


def _render_tile_sharded_operators(gaussians, camera_params, config, use_sh, shard):
  """Operator-chain form of render_tile_sharded (any dtype / tile size): every rank maps all overlaps, empties the
  ranges of foreign tiles, and an identity-forward / all-reduce-backward node sums the 2D gradients."""
  from .mapper.tile_mapper import map_to_tiles
  from .perspective.projection import apply_with_ndc, camera_position
  from .rasterizer.function import rasterize_with_tiles
  from .rendering import RenderedPoints, Rendering
  from .spherical_harmonics import evaluate_sh_at

  g2d, depths, indexes, ndc = apply_with_ndc(
      *gaussians.shape_tensors(), camera_params.T_camera_world, camera_params.projection, camera_params.image_size,
      camera_params.depth_range, config.blur_cov, config.clamp_margin, config.alpha_threshold)
  if use_sh:
    features = evaluate_sh_at(gaussians.feature, gaussians.position.detach(), indexes,
                              camera_position(camera_params.T_camera_world), unique_indexes=True)
  else:
    features = gaussians.feature[indexes]
  overlap_to_point, tile_ranges = map_to_tiles(g2d, ndc, camera_params.image_size, config)
  lo, hi = shard.tile_range(tile_ranges.shape[0] * tile_ranges.shape[1])
  local_ranges = mask_tile_ranges(tile_ranges, lo, hi)
  g2d_r, features_r = reduce_across_ranks(g2d, features, group=shard.group)
  raster = rasterize_with_tiles(g2d_r, features_r, overlap_to_point, local_ranges.view(-1, 2),
                                camera_params.image_size, config)
  points = RenderedPoints(idx=indexes, depths=depths, gaussians2d=g2d,
                          _visibility=raster.visibility if config.compute_visibility else None,
                          _prune_cost=raster.point_heuristic[:, 0] if config.compute_point_heuristic else None,
                          _split_score=raster.point_heuristic[:, 1] if config.compute_point_heuristic else None,
                          features=features)
  return Rendering(image=raster.image, image_weight=raster.image_weight, points=points, camera=camera_params,
                   config=config), (lo, hi)
