"""ctypes binding of libgsplat_b200.so (C ABI in include/gsplat_b200.h).

The CUDA library IS the product: there is no CPU or pure-torch fallback.  Importing this module without
the built library raises; calling an operator with non-CUDA tensors raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_double, c_int32, c_int64, c_size_t, c_void_p
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / f"libgsplat_b200{os.environ.get('GS_BUILD_VARIANT', '')}.so"

P, I32, I64, D, SZ = c_void_p, c_int32, c_int64, c_double, c_size_t



class RasterConfigC(ctypes.Structure):
  """struct gs_raster_config"""
  _fields_ = [
      ("tile_size", c_int32), ("pixel_stride_x", c_int32), ("pixel_stride_y", c_int32),
      ("antialias", c_int32), ("use_alpha_blending", c_int32), ("compute_visibility", c_int32),
      ("compute_point_heuristic", c_int32), ("reserved", c_int32),
      ("clamp_max_alpha", c_double), ("alpha_threshold", c_double), ("saturate_threshold", c_double),
      ("forward_saturate_eps", c_double),
  ]


class RenderArgsC(ctypes.Structure):
  """struct gs_render_args"""
  _fields_ = [
      ("position", P), ("log_scaling", P), ("rotation", P), ("alpha_logit", P), ("feature", P),
      ("T_camera_world", P), ("projection", P),
      ("n", I64), ("width", I32), ("height", I32),
      ("near_plane", D), ("far_plane", D), ("blur_cov", D), ("clamp_margin", D), ("median_threshold", D),
      ("use_sh", I32), ("sh_degree", I32), ("channels", I32), ("use_depth16", I32), ("want_median", I32),
      ("ordering", I32),
      ("config", RasterConfigC),
      ("points", P), ("depths", P), ("ndc", P), ("indexes", P), ("features", P), ("digest", P),
      ("visibility", P), ("heuristic", P), ("camera_pos", P), ("order", P), ("counts", P), ("cum", P),
      ("ws_project", P), ("ws_project_bytes", SZ), ("ws_order", P), ("ws_order_bytes", SZ),
      ("ws_scan", P), ("ws_scan_bytes", SZ),
      ("image", P), ("image_alpha", P), ("median_image", P), ("tile_ranges", P),
      ("ev_raster_start", P), ("ev_raster_end", P),
      ("tile_counts", P), ("tile_cursor", P), ("tile_totals", P),
      ("records", P), ("flush_records", P), ("hits", P), ("tile_lo", I32), ("tile_hi", I32),
  ]


class RenderBwdArgsC(ctypes.Structure):
  """struct gs_render_bwd_args"""
  _fields_ = [
      ("position", P), ("log_scaling", P), ("rotation", P), ("alpha_logit", P), ("feature", P),
      ("T_camera_world", P), ("projection", P),
      ("n", I64), ("v", I64), ("k", I64), ("width", I32), ("height", I32),
      ("blur_cov", D), ("clamp_margin", D),
      ("use_sh", I32), ("sh_degree", I32), ("channels", I32), ("d_image_strided", I32),
      ("config", RasterConfigC),
      ("indexes", P), ("features", P), ("image", P), ("camera_pos", P), ("digest", P),
      ("overlap_to_point", P), ("tile_ranges", P), ("records", P), ("flush_records", P),
      ("d_image", P), ("d_depths", P),
      ("grad_points", P), ("grad_features", P), ("grad_points_preset", I32), ("grad_features_preset", I32),
      ("heuristic", P),
      ("d_position", P), ("d_log_scaling", P), ("d_rotation", P), ("d_alpha_logit", P),
      ("d_T_camera_world", P), ("d_projection", P), ("d_feature", P),
      ("ev_raster_start", P), ("ev_raster_end", P),
      ("d_image_strides", I64 * 3),
      ("d_camera_pos", P), ("phases", I32),
  ]


_PROJECT_CULL = [P, P, P, P, P, P, I64, I32, I32, D, D, D, D, D, P, SZ, P, P]
_PROJECT_COMPACT = [P, P, P, P, P, P, I64, I32, I32, D, D, D, D, D, P, SZ, P, P, P, P, P, P]
_PROJECT_WRITE = [P, P, P, P, P, P, I64, I32, I32, D, D, D, D, P, P, P, P, P, P]
_PROJECT_BWD = [P, P, P, P, P, P, P, I64, I32, I32, D, D, P, P, P, P, P, P, P, P, P]
_SH_FWD = [P, P, P, P, I64, I32, I32, P, P]
_SH_BWD = [P, P, P, P, P, P, I64, I32, I32, I32, P, P, P, P]
_RASTER_FWD = [P, P, P, P, I64, I64, I32, I32, I32, POINTER(RasterConfigC), P, P, P, P]
_RASTER_BWD = [P, P, P, P, P, P, I64, I64, I32, I32, I32, POINTER(RasterConfigC), P, P, P, P]

SIGNATURES = {
    "gs_version": ([], c_int32),
    "gs_last_error_string": ([], ctypes.c_char_p),
    "gs_project_workspace_bytes": ([I64, POINTER(SZ)], c_int32),
    "gs_project_cull_f32": (_PROJECT_CULL, c_int32), "gs_project_cull_f64": (_PROJECT_CULL, c_int32),
    "gs_project_compact_workspace_bytes": ([I64, I32, POINTER(SZ)], c_int32),
    "gs_project_compact_f32": (_PROJECT_COMPACT, c_int32), "gs_project_compact_f64": (_PROJECT_COMPACT, c_int32),
    "gs_project_write_f32": (_PROJECT_WRITE, c_int32), "gs_project_write_f64": (_PROJECT_WRITE, c_int32),
    "gs_project_bwd_f32": (_PROJECT_BWD, c_int32), "gs_project_bwd_f64": (_PROJECT_BWD, c_int32),
    "gs_camera_position_f32": ([P, P, P], c_int32), "gs_camera_position_f64": ([P, P, P], c_int32),
    "gs_sh_bwd_views_f32": ([P, P, P, I64, I32, I32, I64, I32, P, P], c_int32),
    "gs_sh_pack_factors_f32": ([P, P, P, P, I64, I32, I64, P, P], c_int32),
    "gs_allreduce_peers_f32": ([POINTER(ctypes.c_uint64), I32, I32, I64, P], c_int32),
    "gs_sh_pack_factors_peers_f32": ([P, P, P, P, I64, I32, I64, POINTER(ctypes.c_uint64), I32, I64, P], c_int32),
    "gs_sh_fwd_f32": (_SH_FWD, c_int32), "gs_sh_fwd_f64": (_SH_FWD, c_int32),
    "gs_sh_bwd_f32": (_SH_BWD, c_int32), "gs_sh_bwd_f64": (_SH_BWD, c_int32),
    "gs_tile_count": ([P, I64, I32, I32, I32, D, P, P], c_int32),
    "gs_tile_scan_workspace_bytes": ([I64, POINTER(SZ)], c_int32),
    "gs_tile_scan": ([P, I64, P, P, SZ, P, P], c_int32),
    "gs_tile_emit_keys": ([P, P, P, I64, I32, I32, I32, D, I32, P, P, P], c_int32),
    "gs_sort_pairs_workspace_bytes": ([I64, I32, POINTER(SZ)], c_int32),
    "gs_sort_pairs": ([P, P, P, P, I64, I32, I32, I32, P, SZ, P], c_int32),
    "gs_tile_ranges": ([P, I64, I32, P, I64, P], c_int32),
    "gs_depth_order_workspace_bytes": ([I64, POINTER(SZ)], c_int32),
    "gs_depth_order": ([P, I64, I32, P, P, SZ, P], c_int32),
    "gs_tile_count_ordered": ([P, P, I64, I32, I32, I32, D, P, P], c_int32),
    "gs_tile_emit_ordered": ([P, P, P, I64, I32, I32, I32, D, P, P, P], c_int32),
    "gs_tile_ranges_from_tiles": ([P, I64, P, I64, P], c_int32),
    "gs_tile_count_ordered_hits": ([P, P, I64, I32, I32, I32, D, I32, I32, P, P, P], c_int32),
    "gs_tile_emit_hits": ([P, P, P, P, I64, I32, I32, I32, D, I32, I32, P, P, P], c_int32),
    "gs_tile_bin_count": ([P, I64, I32, I32, I32, D, P, P], c_int32),
    "gs_tile_bin_offsets": ([P, I64, P, P, P, P, P], c_int32),
    "gs_tile_bin_emit": ([P, P, I64, I32, I32, I32, D, I32, P, P, P], c_int32),
    "gs_tile_bin_max_per_tile": ([], c_int32),
    "gs_tile_bin_sort": ([P, P, I64, I32, P, P], c_int32),
    "gs_raster_fwd_f32": (_RASTER_FWD, c_int32), "gs_raster_fwd_f64": (_RASTER_FWD, c_int32),
    "gs_raster_fwd_median_f32": ([P, P, P, P, P, I64, I64, I32, I32, I32, POINTER(RasterConfigC), D, P, P, P, P, P], c_int32),
    "gs_raster_bwd_f32": (_RASTER_BWD, c_int32), "gs_raster_bwd_f64": (_RASTER_BWD, c_int32),
    "gs_raster_digest_bytes": ([I64, POINTER(SZ)], c_int32),
    "gs_raster_digest_f32": ([P, P, P, I64, I32, POINTER(RasterConfigC), P, P], c_int32),
    "gs_raster_fwd_digest_f32": ([P, P, P, I64, I64, I32, I32, I32, POINTER(RasterConfigC), D, P, P, P, P, P], c_int32),
    "gs_raster_bwd_digest_f32": ([P, P, P, P, P, I64, I64, I32, I32, I32, POINTER(RasterConfigC), P, P, P, P], c_int32),
    "gs_raster_bwd_digest_strided_f32": ([P, P, P, P, P, POINTER(I64), I64, I64, I32, I32, I32, POINTER(RasterConfigC), P, P, P, P], c_int32),
    "gs_raster_pack_bytes": ([I64, I32, POINTER(SZ), POINTER(SZ)], c_int32),
    "gs_raster_pack_f32": ([P, P, P, I64, I32, I32, I32, P, P, P], c_int32),
    "gs_raster_pack_sorted_f32": ([P, P, P, I64, I32, I32, I32, P, P, P, P], c_int32),
    "gs_raster_fwd_packed_f32": ([P, P, P, I64, I64, I32, I32, I32, POINTER(RasterConfigC), D, P, P, P, P, P], c_int32),
    "gs_raster_bwd_packed_f32": ([P, P, P, P, P, P, POINTER(I64), I64, I64, I32, I32, I32, POINTER(RasterConfigC), P, P, P, P], c_int32),
    "gs_render_stage_a_f32": ([POINTER(RenderArgsC), POINTER(I64), POINTER(I64), POINTER(I64), P], c_int32),
    "gs_render_stage_b_f32": ([POINTER(RenderArgsC), I64, I64, I64, I64, P, P, P, SZ, P], c_int32),
    "gs_render_forward_f32": ([POINTER(RenderArgsC), I64, P, P, P, SZ, POINTER(I64), POINTER(I64), POINTER(I64),
                               POINTER(I32), P], c_int32),
    "gs_render_backward_f32": ([POINTER(RenderBwdArgsC), P], c_int32),
    "gs_morton_codes64": ([P, I64, POINTER(ctypes.c_float), POINTER(ctypes.c_float), I64, P, P, P], c_int32),
    "gs_optim_step_f32": ([I32, I32, I32, P, P, P, D, I64, I32, P, P, P, P, D, D, D, D, P, P, D, P, P, P], c_int32),
    "gs_optim_update_visibility_f32": ([P, P, P, P, D, D, I64, P, P], c_int32),
}

GS_BWD_RASTER, GS_BWD_FEATURE, GS_BWD_PROJECT = 1, 2, 4   # gs_render_bwd_args.phases

_lib = None


def load() -> ctypes.CDLL:
  """Loads the library (never builds it: run `python -m taichi_splatting_b200.build` / __graft_entry__.build())."""
  global _lib
  if _lib is None:
    if not LIB_PATH.exists():
      raise RuntimeError(
          f"{LIB_PATH} is missing: the CUDA extension is the only implementation of this package "
          "(no CPU fallback). Build it with `python -m taichi_splatting_b200.build`.")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (argtypes, restype) in SIGNATURES.items():
      fn = getattr(lib, name)
      fn.argtypes = argtypes
      fn.restype = restype
    _lib = lib
  return _lib


def check(code: int, what: str) -> None:
  if code != 0:
    msg = load().gs_last_error_string().decode()
    if code == -1:
      raise ValueError(f"{what}: {msg}")
    if code == -2:
      raise NotImplementedError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg} (code {code})")


# Hand-written kernels each entry point launches (CUB scan / onesweep launches are not counted).
OWN_KERNELS = {
    "gs_project_cull_f32": 1, "gs_project_cull_f64": 1, "gs_project_compact_f32": 1, "gs_project_compact_f64": 1, "gs_project_write_f32": 1, "gs_project_write_f64": 1,
    "gs_project_bwd_f32": 1, "gs_project_bwd_f64": 1, "gs_camera_position_f32": 1, "gs_camera_position_f64": 1, "gs_sh_fwd_f32": 1, "gs_sh_fwd_f64": 1,
    "gs_sh_bwd_f32": 1, "gs_sh_bwd_f64": 1, "gs_sh_bwd_views_f32": 1, "gs_sh_pack_factors_f32": 1, "gs_sh_pack_factors_peers_f32": 1, "gs_allreduce_peers_f32": 1, "gs_tile_count": 1, "gs_tile_scan": 1, "gs_tile_emit_keys": 1,
    "gs_tile_ranges": 1, "gs_depth_order": 1, "gs_tile_count_ordered": 1, "gs_tile_emit_ordered": 1, "gs_tile_count_ordered_hits": 1, "gs_tile_emit_hits": 1,
    "gs_tile_ranges_from_tiles": 1, "gs_tile_bin_count": 1, "gs_tile_bin_offsets": 1, "gs_tile_bin_emit": 1,
    "gs_tile_bin_sort": 1, "gs_raster_fwd_f32": 3, "gs_raster_fwd_f64": 1, "gs_raster_fwd_median_f32": 3, "gs_raster_bwd_f32": 3,
    "gs_raster_bwd_f64": 1, "gs_raster_digest_f32": 1, "gs_raster_fwd_digest_f32": 2, "gs_raster_bwd_digest_f32": 2, "gs_raster_bwd_digest_strided_f32": 2,
    "gs_raster_pack_f32": 1, "gs_raster_pack_sorted_f32": 1, "gs_raster_fwd_packed_f32": 1, "gs_raster_bwd_packed_f32": 1,
    # whole-frame drivers: project+compact (our projection functor inside cub::DeviceSelect), V publish, camera position,
    # SH, digest, depth key, count, scan tail (+ K publish) | emit, pack (+ ranges), raster forward | raster backward,
    # projection backward, SH backward
    "gs_render_stage_a_f32": 8, "gs_render_stage_b_f32": 3, "gs_render_backward_f32": 3, "gs_render_forward_f32": 8,
    "gs_optim_step_f32": 1, "gs_optim_update_visibility_f32": 1, "gs_morton_codes64": 1,
}


class Profiler:
  """Optional per-entry-point CUDA-event timing on the launching stream (used by bench.py only)."""

  def __init__(self, only=None):
    self.only = only          # None: every entry point; else a set of names
    self.records = []         # (name, start_event, end_event)
    self.launches = 0

  def stage_ms(self):
    out = {}
    for name, a, b in self.records:
      out.setdefault(name, []).append(a.elapsed_time(b))
    return out


profiler = None   # set to a Profiler() to enable


_tls = threading.local()   # .device: index of the device whose stream the calling thread last asked for (stream_ptr)


def call(name: str, *args, on=None) -> None:
  """Calls a C-ABI entry point.  `on`: the torch stream the call was enqueued on when it is not the current one
  (only used to place the profiler's events).

  The library launches on the CUDA *current* device (kernel attributes, scratch and auxiliary streams are per
  device), so when the tensors of this call live on another device than the thread's current one -- a single
  process driving several GPUs -- the call runs inside `torch.cuda.device(that device)`.  The device is the one the
  caller last passed to `stream_ptr`, which every operator evaluates for the call's stream argument."""
  dev = getattr(_tls, "device", None)
  if dev is not None and dev != torch.cuda.current_device():
    with torch.cuda.device(dev):
      _call(name, args, on)
    return
  _call(name, args, on)


def _call(name: str, args, on) -> None:
  fn = getattr(load(), name)
  prof = profiler
  if prof is None:
    check(fn(*args), name)
    return
  prof.launches += OWN_KERNELS.get(name, 0)
  if prof.only is not None and name not in prof.only:
    check(fn(*args), name)
    return
  stream = on if on is not None else torch.cuda.current_stream()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record(stream)
  check(fn(*args), name)
  b.record(stream)
  prof.records.append((name, a, b))


def suffix(dtype: torch.dtype) -> str:
  if dtype == torch.float32:
    return "f32"
  if dtype == torch.float64:
    return "f64"
  raise NotImplementedError(f"dtype {dtype} not supported (float32, float64)")


def require_cuda(**tensors) -> None:
  """Reference convention: `assert arg.is_cuda` (cuda_lib/__init__.py:12-13).  There is no CPU path."""
  for name, t in tensors.items():
    assert t.is_cuda, f"{name}: device must be a CUDA device, got {t.device} (this package has no CPU path)"


def ptr(t) -> int | None:
  """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
  if t is None:
    return None
  assert t.is_cuda, f"expected a CUDA tensor, got device {t.device} (this package has no CPU path)"
  assert t.is_contiguous(), "expected a contiguous tensor"
  return t.data_ptr()


def stream_ptr(device) -> int:
  """Handle of the caller's current torch stream on `device`; also tells `call` which device the call belongs to."""
  device = torch.device(device)
  _tls.device = device.index if device.index is not None else torch.cuda.current_device()
  return torch.cuda.current_stream(device).cuda_stream


def workspace(nbytes: int, device) -> torch.Tensor:
  return torch.empty((max(int(nbytes), 1),), dtype=torch.uint8, device=device)


_host_words = {}


def host_word(device) -> torch.Tensor:
  """A pinned int32 word per device for the asynchronous V / K read-backs."""
  key = (device.type, device.index)
  if key not in _host_words:
    _host_words[key] = torch.zeros((1,), dtype=torch.int32).pin_memory()
  return _host_words[key]


def host_words(device, count: int) -> torch.Tensor:
  """`count` pinned int32 words per device (asynchronous multi-word read-backs)."""
  key = (device.type, device.index, count)
  if key not in _host_words:
    _host_words[key] = torch.zeros((count,), dtype=torch.int32).pin_memory()
  return _host_words[key]


def read_host_word(word: torch.Tensor, device) -> int:
  torch.cuda.current_stream(device).synchronize()
  return int(word.item())


def raster_config_c(config, forward_saturate_eps=None) -> RasterConfigC:
  eps = getattr(config, "forward_saturate_eps", 0.0) if forward_saturate_eps is None else forward_saturate_eps
  return RasterConfigC(
      int(config.tile_size), int(config.pixel_stride[0]), int(config.pixel_stride[1]), int(config.antialias),
      int(config.use_alpha_blending), int(config.compute_visibility), int(config.compute_point_heuristic), 0,
      float(config.clamp_max_alpha), float(config.alpha_threshold), float(config.saturate_threshold), float(eps))
