"""ctypes binding of the C oracle (oracle/gs_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module; the product package never does.

Every function takes and returns numpy arrays (or torch CPU tensors, converted).
Reference anchors are given per function; the C file cites them line by line.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIBS = {}


def build(force: bool = False) -> None:
  """Compile liboracle_f32.so / liboracle_f64.so with the committed Makefile."""
  targets = [_HERE / "liboracle_f32.so", _HERE / "liboracle_f64.so"]
  src = _HERE / "gs_oracle.c"
  stale = any((not t.exists()) or t.stat().st_mtime < src.stat().st_mtime for t in targets)
  if force or stale:
    subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True,
                   stdout=subprocess.DEVNULL)


class _Cfg(ctypes.Structure):
  _fields_ = [
      ("tile_size", ctypes.c_int),
      ("pixel_stride_x", ctypes.c_int),
      ("pixel_stride_y", ctypes.c_int),
      ("antialias", ctypes.c_int),
      ("use_alpha_blending", ctypes.c_int),
      ("compute_visibility", ctypes.c_int),
      ("compute_point_heuristic", ctypes.c_int),
      ("emulate_stale_group", ctypes.c_int),
      ("clamp_max_alpha", ctypes.c_double),
      ("alpha_threshold", ctypes.c_double),
      ("saturate_threshold", ctypes.c_double),
  ]


@dataclass(frozen=True)
class OracleConfig:
  """Field names follow the reference RasterConfig (data_types.py:16-46)."""
  tile_size: int = 16
  pixel_stride: Tuple[int, int] = (2, 2)
  clamp_margin: float = 0.15
  antialias: bool = False
  blur_cov: float = 0.3
  clamp_max_alpha: float = 0.99
  alpha_threshold: float = 1.0 / 255.0
  saturate_threshold: float = 0.9999
  use_alpha_blending: bool = True
  compute_point_heuristic: bool = False
  compute_visibility: bool = False
  median_threshold: float = 0.25


def _cfg(config, emulate_stale_group=False) -> _Cfg:
  return _Cfg(int(config.tile_size), int(config.pixel_stride[0]), int(config.pixel_stride[1]),
              int(config.antialias), int(config.use_alpha_blending),
              int(config.compute_visibility), int(config.compute_point_heuristic),
              int(emulate_stale_group), float(config.clamp_max_alpha),
              float(config.alpha_threshold), float(config.saturate_threshold))


def _lib(dtype) -> ctypes.CDLL:
  key = "f64" if np.dtype(dtype) == np.float64 else "f32"
  if key not in _LIBS:
    build()
    _LIBS[key] = ctypes.CDLL(str(_HERE / f"liboracle_{key}.so"))
  return _LIBS[key]


def _np(x, dtype=None):
  if hasattr(x, "detach"):
    x = x.detach().cpu().numpy()
  x = np.ascontiguousarray(x)
  if dtype is not None and x.dtype != dtype:
    x = x.astype(dtype)
  return x


def _p(a: Optional[np.ndarray]):
  return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def pad_to_tile(image_size, tile_size):
  """mapper/tile_mapper.py:20-24"""
  return tuple(int(math.ceil(x / tile_size) * tile_size) for x in image_size)


# ----------------------------------------------------------------------------- tile mapper
def tile_counts(gaussians, image_size, config) -> np.ndarray:
  """tile_overlaps_kernel, mapper/tile_mapper.py:75-86 (image_size is padded here)."""
  g = _np(gaussians, np.float32)
  w_pad, h_pad = pad_to_tile(image_size, config.tile_size)
  counts = np.empty((g.shape[0],), np.int32)
  _lib(np.float32).orc_tile_counts(_p(g), ctypes.c_int64(g.shape[0]), w_pad, h_pad,
                                   int(config.tile_size),
                                   ctypes.c_float(config.alpha_threshold), _p(counts))
  return counts


def full_cumsum(counts) -> Tuple[np.ndarray, int]:
  """cuda_lib/__init__.py:16-25, full_cumsum.cu:16-47"""
  c = _np(counts, np.int32)
  out = np.empty((c.shape[0] + 1,), np.int32)
  lib = _lib(np.float32)
  lib.orc_full_cumsum.restype = ctypes.c_int64
  total = lib.orc_full_cumsum(_p(c), ctypes.c_int64(c.shape[0]), _p(out))
  return out, int(total)


def tile_keys(gaussians, depths, cum, total, image_size, config, use_depth16=False):
  """generate_sort_keys_kernel, mapper/tile_mapper.py:114-146"""
  g = _np(gaussians, np.float32)
  d = _np(depths, np.float32).reshape(-1)
  cum = _np(cum, np.int32)
  w_pad, h_pad = pad_to_tile(image_size, config.tile_size)
  keys = np.empty((total,), np.uint64)
  o2p = np.empty((total,), np.int32)
  _lib(np.float32).orc_tile_keys(_p(g), _p(d), _p(cum), ctypes.c_int64(g.shape[0]), w_pad, h_pad,
                                 int(config.tile_size), ctypes.c_float(config.alpha_threshold),
                                 int(use_depth16), _p(keys), _p(o2p))
  return keys, o2p


def sort_pairs(keys, values, end_bit=48):
  """cuda_lib.radix_sort_pairs (stable LSD on bits [0,end_bit)), radix_sort_pairs.cu:7-29"""
  k = _np(keys, np.uint64)
  v = _np(values, np.int32)
  ko, vo = np.empty_like(k), np.empty_like(v)
  _lib(np.float32).orc_sort_pairs(_p(k), _p(v), ctypes.c_int64(k.shape[0]), int(end_bit),
                                  _p(ko), _p(vo))
  return ko, vo


def tile_ranges(sorted_keys, num_tiles, use_depth16=False) -> np.ndarray:
  """find_ranges_kernel, mapper/tile_mapper.py:92-112"""
  k = _np(sorted_keys, np.uint64)
  ranges = np.zeros((num_tiles, 2), np.int32)
  _lib(np.float32).orc_tile_ranges(_p(k), ctypes.c_int64(k.shape[0]), int(use_depth16), _p(ranges))
  return ranges


def map_to_tiles(gaussians, depth, image_size, config, use_depth16=False, return_keys=False):
  """mapper/tile_mapper.py:171-198 -> (overlap_to_point (K,), tile_ranges (TH,TW,2))."""
  ts = config.tile_size
  w_pad, h_pad = pad_to_tile(image_size, ts)
  tile_shape = (h_pad // ts, w_pad // ts)
  assert tile_shape[0] * tile_shape[1] < 65535, "tile count exceeds 16 bit id"
  counts = tile_counts(gaussians, image_size, config)
  cum, total = full_cumsum(counts)
  if total > 0:
    keys, o2p = tile_keys(gaussians, depth, cum[:-1], total, image_size, config, use_depth16)
    skeys, so2p = sort_pairs(keys, o2p, end_bit=32 if use_depth16 else 48)
    ranges = tile_ranges(skeys, tile_shape[0] * tile_shape[1], use_depth16)
  else:
    skeys, so2p = np.empty((0,), np.uint64), np.empty((0,), np.int32)
    ranges = np.zeros((tile_shape[0] * tile_shape[1], 2), np.int32)
  ranges = ranges.reshape(*tile_shape, 2)
  if return_keys:
    return so2p, ranges, skeys, counts
  return so2p, ranges


# ----------------------------------------------------------------------------- rasteriser
def raster_forward(points, features, tile_ranges_, overlap_to_point, image_size, config,
                   dtype=np.float32, emulate_stale_group=False):
  """_forward_kernel, rasterizer/forward.py:22-135 -> (image (H,W,F), alpha (H,W), visibility (V,)|None)"""
  pts = _np(points, dtype)
  feat = _np(features, dtype)
  rng = _np(tile_ranges_, np.int32).reshape(-1, 2)
  o2p = _np(overlap_to_point, np.int32)
  w, h = int(image_size[0]), int(image_size[1])
  F = feat.shape[1]
  assert F <= 16
  image = np.zeros((h, w, F), dtype)
  alpha = np.zeros((h, w), dtype)
  vis = np.zeros((pts.shape[0],), dtype) if config.compute_visibility else None
  cfg = _cfg(config, emulate_stale_group)
  suffix = "f64" if np.dtype(dtype) == np.float64 else "f32"
  getattr(_lib(dtype), f"orc_raster_fwd_{suffix}")(
      _p(pts), _p(feat), _p(rng), _p(o2p), w, h, F, ctypes.byref(cfg), _p(image), _p(alpha), _p(vis))
  return image, alpha, vis


def raster_backward(points, features, tile_ranges_, overlap_to_point, image, grad_image, image_size,
                    config, dtype=np.float32, emulate_stale_group=False):
  """_backward_kernel, rasterizer/backward.py:50-225 -> (grad_points (V,7), grad_features (V,F), heuristic (V,2)|None)"""
  pts = _np(points, dtype)
  feat = _np(features, dtype)
  rng = _np(tile_ranges_, np.int32).reshape(-1, 2)
  o2p = _np(overlap_to_point, np.int32)
  img = _np(image, dtype)
  gimg = _np(grad_image, dtype)
  w, h = int(image_size[0]), int(image_size[1])
  F = feat.shape[1]
  assert F <= 16
  gp = np.zeros_like(pts)
  gf = np.zeros_like(feat)
  heur = np.zeros((pts.shape[0], 2), dtype) if config.compute_point_heuristic else None
  cfg = _cfg(config, emulate_stale_group)
  suffix = "f64" if np.dtype(dtype) == np.float64 else "f32"
  getattr(_lib(dtype), f"orc_raster_bwd_{suffix}")(
      _p(pts), _p(feat), _p(rng), _p(o2p), _p(img), _p(gimg), w, h, F, ctypes.byref(cfg),
      _p(gp), _p(gf), _p(heur))
  return gp, gf, heur


def num_threads() -> int:
  return int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
