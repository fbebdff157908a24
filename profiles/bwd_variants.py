"""What the raster backward costs with parts of its outputs switched off (bench workload, operator chain): the full
kernel (packed-2D gradient + feature gradient + heuristics) against instantiations without the feature-gradient
panel plane and / or without the heuristics.  An upper bound for what restructuring the panel could buy.
Usage: python profiles/bwd_variants.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import taichi_splatting_b200 as ts
from taichi_splatting_b200 import _lib
from taichi_splatting_b200.benchmarks import scenes
from taichi_splatting_b200.mapper.tile_mapper import map_to_tiles
from taichi_splatting_b200.perspective.projection import apply_with_ndc, camera_position
from taichi_splatting_b200.rasterizer.function import rasterize_with_tiles
from taichi_splatting_b200.spherical_harmonics import evaluate_sh_at

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
cam = scenes.benchmark_camera((2048, 2048))
cloud = scenes.random_3d_gaussians(1_000_000, cam, sh_degree=3, seed=0).to(dev)
camera = cam.to(device=dev)
with torch.no_grad():
  cfg0 = ts.RasterConfig()
  g2d, depths, indexes, ndc = apply_with_ndc(*cloud.shape_tensors(), camera.T_camera_world, camera.projection,
                                             camera.image_size, camera.depth_range, cfg0.blur_cov, cfg0.clamp_margin,
                                             cfg0.alpha_threshold)
  feats = evaluate_sh_at(cloud.feature, cloud.position, indexes, camera_position(camera.T_camera_world), unique_indexes=True)
  o2p, ranges = map_to_tiles(g2d, ndc, image_size=camera.image_size, config=cfg0)
  ranges = ranges.view(-1, 2)
print(f"V = {g2d.shape[0]}, K = {o2p.shape[0]}")
for need_g, need_f, heur in ((True, True, True), (True, False, True), (True, True, False), (True, False, False),
                             (False, True, False)):
  config = ts.RasterConfig(compute_visibility=True, compute_point_heuristic=heur)
  g = g2d.clone().requires_grad_(need_g)
  f = feats.clone().requires_grad_(need_f)
  prof = _lib.Profiler(only={"gs_raster_bwd_packed_f32", "gs_raster_fwd_packed_f32"})
  for it in range(steps + 3):
    if it == 3:
      _lib.profiler = prof
    out = rasterize_with_tiles(g, f, o2p, ranges, camera.image_size, config)
    out.image.sum().backward()
  torch.cuda.synchronize()
  _lib.profiler = None
  line = []
  for k, v in prof.stage_ms().items():
    v = sorted(v)
    line.append(f"{k} median {v[len(v) // 2]:.4f} min {v[0]:.4f}")
  print(f"grad points {need_g!s:5} grad features {need_f!s:5} heuristics {heur!s:5}: " + "; ".join(line))
