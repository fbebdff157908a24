"""Tile mapper operator (reference: taichi_splatting/mapper/tile_mapper.py:20-224).

count -> scan -> (tile|depth) keys -> radix sort -> ranges; device work in csrc/mapper.cu.  One host read
(K, the total overlap count) per call, stream-scoped (the reference does two device-wide syncs, D11).
"""
import math
import os
from numbers import Integral
from beartype.typing import Tuple

import torch
from beartype import beartype

from .. import _lib
from ..data_types import RasterConfig

MAX_TILES = 65535


# Which of the two orderings with the reference's result the mapper runs by default.  "two_level": stable radix
# sorts (depth on V Gaussians, tile on K overlaps), linear in K for any scene.  "binned": per-tile atomics and one
# shared-memory bitonic sort per tile -- no global sort, but O(n log^2 n) in the tile population: on par at ~250
# overlaps per tile (bench workload: 2.21 vs 2.22 ms per step), far behind on crowded tiles.  Hence the default.
ORDERING = os.environ.get("GS_ORDERING", "two_level")


def pad_to_tile(image_size: Tuple[Integral, Integral], tile_size: int):
  return tuple(int(math.ceil(x / tile_size) * tile_size) for x in image_size)


def key_bits(num_tiles: int, use_depth16: bool) -> int:
  """Radix-sort bit range.  The reference sorts 48 (or 32) bits; bits above ceil(log2(T)) are zero for every
  key, so sorting only the populated bits yields the identical order with fewer onesweep passes."""
  tile_bits = max(1, (max(num_tiles, 1) - 1).bit_length())
  return (16 if use_depth16 else 32) + tile_bits


@beartype
def map_to_tiles(gaussians: torch.Tensor, depth: torch.Tensor, image_size: Tuple[Integral, Integral],
                 config: RasterConfig, use_depth16: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
  """-> overlap_to_point (K,) int32 sorted by (tile, depth, index); tile_ranges (TH, TW, 2) int32."""
  assert gaussians.ndim == 2 and gaussians.shape[1] == 7, f"gaussians must be Nx7 got {gaussians.shape}"
  assert depth.ndim == 2 and depth.shape[1] == 1, f"depths must be Nx1, got {depth.shape}"
  assert gaussians.shape[0] == depth.shape[0], f"size mismatch {gaussians.shape} vs {depth.shape}"
  _lib.require_cuda(gaussians=gaussians, depth=depth)
  with torch.no_grad():
    g = gaussians.detach().to(torch.float32).contiguous()   # the mapper is f32-only (reference :14)
    d = depth.detach().to(torch.float32).contiguous().view(-1)
    binned = bin_and_sort_binned(g, d, image_size, config, use_depth16) if ORDERING == "binned" else None
    if binned is not None:
      return binned
    o2p, ranges, _, _, _ = bin_and_sort(g, d, image_size, config, use_depth16)
  return o2p, ranges


def tile_bits(num_tiles: int) -> int:
  return max(1, (max(num_tiles, 1) - 1).bit_length())


def bin_and_sort_binned(g: torch.Tensor, d: torch.Tensor, image_size, config, use_depth16: bool = False):
  """The binned ordering on contiguous fp32 inputs g (V,7), d (V,) -> (overlap_to_point (K,), tile_ranges (TH,TW,2)),
  or None when one tile holds more overlaps than the shared-memory sort takes (caller falls back to bin_and_sort).

  Same final order as the reference's 48-bit LSD sort over (tile | depth) with ascending-index ties
  (tile_mapper.py:148-157), reached without a global sort: per-tile counts (atomics) -> tile offsets = tile ranges ->
  every overlap takes a slot in its tile's segment -> each segment sorted on (depth bits, index) in shared memory."""
  device = g.device
  ts = config.tile_size
  w_pad, h_pad = pad_to_tile(image_size, ts)
  tile_shape = (h_pad // ts, w_pad // ts)
  num_tiles = tile_shape[0] * tile_shape[1]
  assert num_tiles < MAX_TILES, \
      f"tile dimensions {tile_shape} for image size {image_size} exceed maximum tile count (16 bit id), try increasing tile_size"
  v = g.shape[0]
  call, ptr = _lib.call, _lib.ptr
  stream = _lib.stream_ptr(device)
  thr = float(config.alpha_threshold)
  tile_counts = torch.empty((num_tiles,), dtype=torch.int32, device=device)
  cursor = torch.empty((num_tiles,), dtype=torch.int32, device=device)
  totals = torch.empty((2,), dtype=torch.int32, device=device)
  tile_ranges = torch.empty((*tile_shape, 2), dtype=torch.int32, device=device)
  words = _lib.host_words(device, 2)
  call("gs_tile_bin_count", ptr(g), v, w_pad, h_pad, ts, thr, ptr(tile_counts), stream)
  call("gs_tile_bin_offsets", ptr(tile_counts), num_tiles, ptr(tile_ranges), ptr(cursor), ptr(totals), words.data_ptr(),
       stream)
  torch.cuda.current_stream(device).synchronize()     # the one host read of the mapper: K and the largest tile
  k, max_per_tile = int(words[0]), int(words[1])
  if max_per_tile > _lib.load().gs_tile_bin_max_per_tile():
    return None
  keys = torch.empty((k,), dtype=torch.int64, device=device)
  o2p = torch.empty((k,), dtype=torch.int32, device=device)
  if k > 0:
    call("gs_tile_bin_emit", ptr(g), ptr(d), v, w_pad, h_pad, ts, thr, int(use_depth16), ptr(cursor), ptr(keys), stream)
    call("gs_tile_bin_sort", ptr(keys), ptr(tile_ranges), num_tiles, max_per_tile, ptr(o2p), stream)
  return o2p, tile_ranges


def bin_and_sort(g: torch.Tensor, d: torch.Tensor, image_size, config, use_depth16: bool = False,
                 tile_range=(0, 0)):
  """The two-level ordering on contiguous fp32 inputs g (V,7), d (V,) -> (overlap_to_point (K,), tile_ranges
  (TH,TW,2), sorted tile ids (K,) int32, order (V,) int32, counts in depth order (V,) int32).

  Same final order as the reference's single 48-bit LSD radix sort over (tile | depth) (tile_mapper.py:148-157):
  such a sort is a stable sort by depth followed by a stable sort by tile, and all overlaps of one Gaussian share
  its depth, so the depth passes run on the V Gaussians before the expansion to K overlaps (gs_depth_order), the
  overlaps are emitted in that order keyed by tile id only, and one stable sort on ceil(log2 T) bits finishes it.
  tile_range (lo, hi): keep only the overlaps of tiles lo <= id < hi (a tile-sharded multi-GPU rank); (0, 0): all."""
  device = g.device
  ts = config.tile_size
  w_pad, h_pad = pad_to_tile(image_size, ts)
  tile_shape = (h_pad // ts, w_pad // ts)
  num_tiles = tile_shape[0] * tile_shape[1]
  assert num_tiles < MAX_TILES, \
      f"tile dimensions {tile_shape} for image size {image_size} exceed maximum tile count (16 bit id), try increasing tile_size"
  v = g.shape[0]
  call, ptr = _lib.call, _lib.ptr
  stream = _lib.stream_ptr(device)
  thr = float(config.alpha_threshold)
  nbytes = _lib.c_size_t()

  order = torch.empty((v,), dtype=torch.int32, device=device)
  call("gs_depth_order_workspace_bytes", v, nbytes)
  ws = _lib.workspace(nbytes.value, device)
  call("gs_depth_order", ptr(d), v, int(use_depth16), ptr(order), ws.data_ptr(), ws.numel(), stream)
  counts = torch.empty((v,), dtype=torch.int32, device=device)
  cum = torch.empty((v + 1,), dtype=torch.int32, device=device)
  # count and emit share ONE grid query: the count kernel leaves a 16-byte hit record per Gaussian for the emit kernel
  hits = torch.empty((v, 2), dtype=torch.int64, device=device)
  call("gs_tile_count_ordered_hits", ptr(g), ptr(order), v, w_pad, h_pad, ts, thr, tile_range[0], tile_range[1], ptr(counts),
       ptr(hits), stream)
  call("gs_tile_scan_workspace_bytes", v, nbytes)
  ws2 = _lib.workspace(nbytes.value, device)
  word = _lib.host_word(device)
  call("gs_tile_scan", ptr(counts), v, ptr(cum), ws2.data_ptr(), ws2.numel(), word.data_ptr(), stream)
  tile_ranges = torch.empty((*tile_shape, 2), dtype=torch.int32, device=device)
  k = _lib.read_host_word(word, device)     # the one host read of the mapper (reference: two device-wide syncs)

  tiles = torch.empty((2, k), dtype=torch.int32, device=device)
  o2p = torch.empty((2, k), dtype=torch.int32, device=device)
  if k > 0:
    call("gs_tile_emit_hits", ptr(g), ptr(order), ptr(cum), ptr(hits), v, w_pad, h_pad, ts, thr, tile_range[0], tile_range[1],
         ptr(tiles[0]), ptr(o2p[0]), stream)
    call("gs_sort_pairs_workspace_bytes", k, 4, nbytes)
    ws3 = _lib.workspace(nbytes.value, device)
    call("gs_sort_pairs", ptr(tiles[0]), ptr(o2p[0]), ptr(tiles[1]), ptr(o2p[1]), k, 4, 0, tile_bits(num_tiles),
         ws3.data_ptr(), ws3.numel(), stream)
  call("gs_tile_ranges_from_tiles", ptr(tiles[1]), k, ptr(tile_ranges), num_tiles, stream)
  return o2p[1], tile_ranges, tiles[1], order, counts


def map_to_tiles_full(gaussians, depth, image_size, config, use_depth16=False, two_level=True, binned=False):
  """As map_to_tiles, also returning the sorted (tile | depth) keys and per-Gaussian counts (tests / diagnostics).
  two_level=False runs the reference's own sequence (count, scan, 64-bit keys, one 48-bit sort, ranges);
  binned=True the per-tile shared-memory sort (keys / counts are then rebuilt from the result)."""
  _lib.require_cuda(gaussians=gaussians, depth=depth)
  if binned:
    with torch.no_grad():
      g = gaussians.detach().to(torch.float32).contiguous()
      d = depth.detach().to(torch.float32).contiguous().view(-1)
      out = bin_and_sort_binned(g, d, image_size, config, use_depth16)
      if out is None:   # a tile exceeds the shared-memory sort capacity: callers fall back to the two-level ordering
        return None
      o2p, tile_ranges = out
      r = tile_ranges.view(-1, 2).long()
      tiles = torch.repeat_interleave(torch.arange(r.shape[0], device=g.device), r[:, 1] - r[:, 0]).to(torch.int32)
      counts = torch.bincount(o2p.long(), minlength=g.shape[0]).to(torch.int32)
      if use_depth16:
        dbits = (d.clamp(0, 1) * 65535.0).to(torch.int32)
        keys = (tiles << 16) | dbits[o2p.long()]
      else:
        keys = (tiles.to(torch.int64) << 32) | (d.view(torch.int32)[o2p.long()].to(torch.int64) & 0xFFFFFFFF)
      return o2p, tile_ranges, keys, counts
  if two_level:
    with torch.no_grad():
      g = gaussians.detach().to(torch.float32).contiguous()
      d = depth.detach().to(torch.float32).contiguous().view(-1)
      o2p, tile_ranges, tiles, order, counts_sorted = bin_and_sort(g, d, image_size, config, use_depth16)
      counts = torch.empty_like(counts_sorted)
      counts[order.long()] = counts_sorted
      if use_depth16:
        dbits = (d.clamp(0, 1) * 65535.0).to(torch.int32)
        keys = (tiles << 16) | dbits[o2p.long()]
      else:
        keys = (tiles.to(torch.int64) << 32) | (d.view(torch.int32)[o2p.long()].to(torch.int64) & 0xFFFFFFFF)
      return o2p, tile_ranges, keys, counts
  return _map_to_tiles_single_sort(gaussians, depth, image_size, config, use_depth16)


def _map_to_tiles_single_sort(gaussians, depth, image_size, config, use_depth16=False):
  device = gaussians.device
  ts = config.tile_size
  w_pad, h_pad = pad_to_tile(image_size, ts)
  tile_shape = (h_pad // ts, w_pad // ts)
  num_tiles = tile_shape[0] * tile_shape[1]
  assert num_tiles < MAX_TILES, \
      f"tile dimensions {tile_shape} for image size {image_size} exceed maximum tile count (16 bit id), try increasing tile_size"

  with torch.no_grad():
    g = gaussians.detach().to(torch.float32).contiguous()   # the mapper is f32-only (reference :14)
    d = depth.detach().to(torch.float32).contiguous().view(-1)
    v = g.shape[0]
    stream = _lib.stream_ptr(device)
    key_dtype, key_bytes = (torch.int32, 4) if use_depth16 else (torch.int64, 8)
    tile_ranges = torch.empty((*tile_shape, 2), dtype=torch.int32, device=device)

    counts = torch.empty((v,), dtype=torch.int32, device=device)
    cum = torch.empty((v + 1,), dtype=torch.int32, device=device)
    _lib.call("gs_tile_count", _lib.ptr(g), v, w_pad, h_pad, ts, float(config.alpha_threshold),
              _lib.ptr(counts), stream)
    nbytes = _lib.c_size_t()
    _lib.call("gs_tile_scan_workspace_bytes", v, nbytes)
    ws = _lib.workspace(nbytes.value, device)
    word = _lib.host_word(device)
    _lib.call("gs_tile_scan", _lib.ptr(counts), v, _lib.ptr(cum), ws.data_ptr(), ws.numel(), word.data_ptr(), stream)
    k = _lib.read_host_word(word, device)

    keys = torch.empty((k,), dtype=key_dtype, device=device)
    o2p = torch.empty((k,), dtype=torch.int32, device=device)
    keys_sorted = torch.empty_like(keys)
    o2p_sorted = torch.empty_like(o2p)
    if k > 0:
      _lib.call("gs_tile_emit_keys", _lib.ptr(g), _lib.ptr(d), _lib.ptr(cum), v, w_pad, h_pad, ts,
                float(config.alpha_threshold), int(use_depth16), _lib.ptr(keys), _lib.ptr(o2p), stream)
      _lib.call("gs_sort_pairs_workspace_bytes", k, key_bytes, nbytes)
      ws = _lib.workspace(nbytes.value, device)
      _lib.call("gs_sort_pairs", _lib.ptr(keys), _lib.ptr(o2p), _lib.ptr(keys_sorted), _lib.ptr(o2p_sorted), k,
                key_bytes, 0, key_bits(num_tiles, use_depth16), ws.data_ptr(), ws.numel(), stream)
    _lib.call("gs_tile_ranges", _lib.ptr(keys_sorted), k, key_bytes, _lib.ptr(tile_ranges), num_tiles, stream)
    return o2p_sorted, tile_ranges, keys_sorted, counts
