"""CPU tests of the host side: C-ABI library loads and exports every declared symbol (no compute calls),
value types, API surface, and that the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  text = open(os.path.join(ROOT, "include", "gsplat_b200.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(gs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
  from taichi_splatting_b200 import _lib
  assert _lib.LIB_PATH.exists(), "libgsplat_b200.so not built: run __graft_entry__.build()"
  lib = ctypes.CDLL(str(_lib.LIB_PATH))
  names = _declared_symbols()
  assert len(names) >= 24
  for n in names:
    assert hasattr(lib, n), f"{n} declared in include/gsplat_b200.h but not exported"
    assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.SIGNATURES"
  assert set(_lib.SIGNATURES) == set(names)
  assert _lib.load().gs_version() >= 100
  # size queries are host-only and safe without a GPU
  nbytes = ctypes.c_size_t()
  _lib.call("gs_project_workspace_bytes", 1000, nbytes)
  assert nbytes.value >= 8000
  _lib.call("gs_sort_pairs_workspace_bytes", 100000, 8, nbytes)
  assert nbytes.value > 0
  assert ctypes.sizeof(_lib.RasterConfigC) == 64


def test_api_surface_matches_reference():
  import taichi_splatting_b200 as ts
  # reference taichi_splatting/__init__.py:1-33
  for name in ["render_gaussians", "Rendering", "map_to_tiles", "pad_to_tile", "Gaussians2D", "Gaussians3D",
               "RasterConfig", "evaluate_sh_at", "rasterize", "rasterize_with_tiles", "perspective", "TaichiQueue"]:
    assert hasattr(ts, name), name
  sig = inspect.signature(ts.render_gaussians)
  assert list(sig.parameters) == ["gaussians", "camera_params", "config", "use_sh", "render_depth", "use_depth16",
                                  "render_median_depth"]
  sig = inspect.signature(ts.perspective.apply)
  assert list(sig.parameters) == ["position", "log_scaling", "rotation", "alpha_logit", "T_camera_world", "projection",
                                  "image_size", "depth_range", "blur_cov", "clamp_margin", "alpha_threshold"]
  assert list(inspect.signature(ts.rasterize_with_tiles).parameters) == [
      "gaussians2d", "features", "overlap_to_point", "tile_overlap_ranges", "image_size", "config"]
  assert list(inspect.signature(ts.map_to_tiles).parameters) == ["gaussians", "depth", "image_size", "config", "use_depth16"]
  assert list(inspect.signature(ts.evaluate_sh_at).parameters)[:4] == ["sh_params", "positions", "indexes", "camera_pos"]
  assert ts.RasterOut._fields == ("image", "image_weight", "point_heuristic", "visibility")


def test_raster_config_and_containers():
  import taichi_splatting_b200 as ts
  from dataclasses import replace
  c = ts.RasterConfig()
  assert (c.tile_size, c.pixel_stride, c.clamp_margin, c.blur_cov, c.clamp_max_alpha, c.saturate_threshold,
          c.median_threshold) == (16, (2, 2), 0.15, 0.3, 0.99, 0.9999, 0.25)
  assert abs(c.alpha_threshold - 1 / 255) < 1e-12 and c.use_alpha_blending and not c.antialias
  assert hash(c) == hash(ts.RasterConfig()) and replace(c, tile_size=8) != c
  with pytest.raises(Exception):
    c.tile_size = 8
  assert ts.pad_to_tile((100, 33), 16) == (112, 48)
  from taichi_splatting_b200.mapper.tile_mapper import key_bits
  assert key_bits(16384, False) == 46 and key_bits(4096, False) == 44 and key_bits(300, True) == 25
  n = 5
  g = ts.Gaussians3D(position=torch.zeros(n, 3), log_scaling=torch.zeros(n, 3), rotation=torch.zeros(n, 4),
                     alpha_logit=torch.zeros(n, 1), feature=torch.zeros(n, 3), batch_size=(n,))
  assert g.batch_size == (n,) and g.packed().shape == (n, 11) and len(g.shape_tensors()) == 4
  assert g[1:3].position.shape == (2, 3) and g.to(torch.float64).position.dtype == torch.float64
  with pytest.raises(AssertionError):
    ts.Gaussians3D(position=torch.zeros(n, 2), log_scaling=torch.zeros(n, 3), rotation=torch.zeros(n, 4),
                   alpha_logit=torch.zeros(n, 1), feature=torch.zeros(n, 3))
  cam = ts.perspective.CameraParams(projection=torch.tensor([100., 100., 32., 24.]), T_camera_world=torch.eye(4),
                                    near_plane=0.1, far_plane=100., image_size=(64, 48))
  assert cam.depth_range == (0.1, 100.) and torch.allclose(cam.camera_position, torch.zeros(3))
  ts.TaichiQueue.init(arch=None)
  assert ts.TaichiQueue.run_sync(lambda a, b: a + b, 1, 2) == 3
  ts.TaichiQueue.stop()


def test_no_cpu_fallback():
  """The operators are CUDA-only: CPU tensors must be rejected, never silently computed elsewhere."""
  import taichi_splatting_b200 as ts
  cfg = ts.RasterConfig()
  with pytest.raises(AssertionError, match="CUDA"):
    ts.map_to_tiles(torch.zeros(4, 7), torch.zeros(4, 1), (64, 64), cfg)
  with pytest.raises(AssertionError, match="CUDA"):
    ts.rasterize_with_tiles(torch.zeros(4, 7), torch.zeros(4, 3), torch.zeros(0, dtype=torch.int32),
                            torch.zeros(16, 2, dtype=torch.int32), (64, 64), cfg)
  with pytest.raises(AssertionError, match="CUDA"):
    ts.evaluate_sh_at(torch.zeros(4, 3, 4), torch.zeros(4, 3), torch.zeros(2, dtype=torch.int64), torch.zeros(3))
  # and the package does not import the oracle
  import sys
  src = "".join(open(os.path.join(dp, f)).read() for dp, _, fs in os.walk(os.path.join(ROOT, "taichi_splatting_b200"))
                for f in fs if f.endswith(".py"))
  assert "import oracle" not in src and "from oracle" not in src


def test_ctypes_structs_match_the_header(tmp_path):
  """The ctypes mirrors of gs_raster_config / gs_render_args / gs_render_bwd_args must have the C compiler's size
  and field offsets (a drifted field silently shifts every pointer after it)."""
  import ctypes
  import subprocess
  from taichi_splatting_b200 import _lib
  structs = {"gs_raster_config": _lib.RasterConfigC, "gs_render_args": _lib.RenderArgsC,
             "gs_render_bwd_args": _lib.RenderBwdArgsC}
  lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gsplat_b200.h"', 'int main(void) {']
  for cname, cls in structs.items():
    lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
    for field, _ in cls._fields_:
      lines.append(f'  printf("{cname}.{field} %zu\\n", offsetof({cname}, {field}));')
  lines += ['  return 0;', '}']
  src = tmp_path / "layout.c"
  src.write_text("\n".join(lines))
  exe = tmp_path / "layout"
  subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
  out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
  for cname, cls in structs.items():
    assert int(out[cname]) == ctypes.sizeof(cls), (cname, out[cname], ctypes.sizeof(cls))
    for field, _ in cls._fields_:
      assert int(out[f"{cname}.{field}"]) == getattr(cls, field).offset, (cname, field)


def test_bench_line_contract():
  """The committed bench line of the final build carries every key the bench contract names, and the algorithmic
  byte model reproduces SURVEY 8d's worked figure for the bench workload (2.22 GB per frame)."""
  import importlib.util
  import json
  line = json.loads(open(os.path.join(ROOT, "profiles", "r02", "r02al_bench.json")).read().strip().splitlines()[-1])
  for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
    assert key in line, key
  assert line["config"]["workload"] and line["gpu_launches"] > 0 and line["higher_is_better"] is True
  for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
    assert key in line["e2e"], key
  for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
    assert key in line["roofline"], key
  for key in ("value", "unit", "cores", "kind", "sample"):
    assert key in line["cpu_baseline"], key
  assert abs(line["value"] - line["n_gpus"] * line["config"]["n_gaussians"] / (line["ms_per_step"] * 1e-3)) < 1e-3 * line["value"]
  spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
  bench = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(bench)
  _, total = bench.algorithmic_bytes(1_000_000, 1_000_000, 3_838_201, 2048 * 2048, 16384, 3, 16, dense_grad_image=True)
  assert abs(total / 1e9 - 2.22) < 0.02, total
  # the bench's image.sum() loss hands the backward an expanded scalar instead of a dense (H,W,3) gradient: 4PF fewer bytes
  _, total_sum = bench.algorithmic_bytes(1_000_000, 1_000_000, 3_838_201, 2048 * 2048, 16384, 3, 16)
  assert total - total_sum == 4 * 2048 * 2048 * 3


def test_no_pointer_is_taken_from_a_temporary():
  """`_lib.ptr(x.contiguous())` hands the library the address of an unnamed tensor that CPython frees as soon as `ptr`
  returns -- before the launch (found by initcheck without the caching allocator, profiles/r02/r02an_initcheck.log).
  Static check over the package: the argument of every `ptr(...)` / `.data_ptr()` is a name, an attribute or a
  subscript of one, never the result of a call."""
  import ast
  import pathlib
  pkg = pathlib.Path(ROOT) / "taichi_splatting_b200"

  def is_stable(node):   # a name, attribute chain or subscript of one keeps its tensor alive; a call result does not
    if isinstance(node, ast.Name):
      return True
    if isinstance(node, ast.Attribute):
      return is_stable(node.value)
    if isinstance(node, ast.Subscript):
      return is_stable(node.value)
    if isinstance(node, ast.Constant):
      return True
    return False

  offenders = []
  for path in sorted(pkg.rglob("*.py")):
    tree = ast.parse(path.read_text())
    for node in ast.walk(tree):
      if not isinstance(node, ast.Call):
        continue
      f = node.func
      name = f.id if isinstance(f, ast.Name) else f.attr if isinstance(f, ast.Attribute) else None
      if name == "ptr" and node.args and not is_stable(node.args[0]):
        if isinstance(node.args[0], ast.IfExp):   # `ptr(a if c else b)` of stable operands is fine
          if all(is_stable(x) for x in (node.args[0].body, node.args[0].orelse)):
            continue
        offenders.append(f"{path.relative_to(ROOT)}:{node.lineno}")
      if name == "data_ptr" and isinstance(f, ast.Attribute) and not is_stable(f.value):
        v = f.value   # `x.untyped_storage().data_ptr()` only compares addresses (parallel.py: adjacency of gradients)
        if isinstance(v, ast.Call) and isinstance(v.func, ast.Attribute) and v.func.attr == "untyped_storage" and is_stable(v.func.value):
          continue
        offenders.append(f"{path.relative_to(ROOT)}:{node.lineno}")
  assert not offenders, offenders
