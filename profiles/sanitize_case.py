"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import taichi_splatting_b200 as ts
from taichi_splatting_b200.benchmarks import scenes

dev = torch.device("cuda:0")
for size, n, ts_, aa in (((200, 136), 6000, 16, False), ((96, 80), 2000, 8, True)):
  cam = scenes.benchmark_camera(size)
  cloud = scenes.random_3d_gaussians(n, cam, scale_factor=2.0, sh_degree=3, seed=1).to(dev).requires_grad_(True)
  cfg = ts.RasterConfig(tile_size=ts_, antialias=aa, compute_visibility=True, compute_point_heuristic=True)
  out = ts.render_gaussians(cloud, cam.to(device=dev), cfg, use_sh=True, render_median_depth=True)
  out.image.sum().backward()
  g2 = scenes.random_2d_gaussians(3000, size, seed=2).to(dev)
  from taichi_splatting_b200.misc.renderer2d import project_gaussians2d
  r = ts.rasterize(project_gaussians2d(g2).requires_grad_(True), g2.depths, g2.feature.requires_grad_(True), size, cfg, use_depth16=True)
  r.image.sum().backward()
torch.cuda.synchronize()
print("ok")
