"""Times render_gaussians forward+backward on the other BASELINE.json configurations (single GPU, device-resident
inputs, CUDA events, median of `steps` after 5 warm-ups).  Usage: python profiles/run_config.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import taichi_splatting_b200 as ts
from taichi_splatting_b200.benchmarks import scenes

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
CONFIGS = [
    ("cfg2: 100 k Gaussians, 1024x1024, SH deg 0 (plain features)", 100_000, (1024, 1024), None, False),
    ("cfg3: 1 M Gaussians, 2048x2048, SH deg 3, vis + heuristics + median depth", 1_000_000, (2048, 2048), 3, True),
    ("cfg4 (one GPU): 6 M Gaussians, 4096x2160, SH deg 3, vis + heuristics + median depth", 6_000_000, (4096, 2160), 3, True),
    ("cfg5 (one view): 1 M Gaussians, 1920x1080, SH deg 3, vis + heuristics + median depth", 1_000_000, (1920, 1080), 3, True),
]
for name, n, size, deg, extras in CONFIGS:
  cam = scenes.benchmark_camera(size)
  cloud = scenes.random_3d_gaussians(n, cam, sh_degree=deg, seed=0).to(dev).requires_grad_(True)
  camera = cam.to(device=dev)
  config = ts.RasterConfig(compute_visibility=extras, compute_point_heuristic=extras)

  def step():
    for t in cloud.to_dict().values():
      t.grad = None
    out = ts.render_gaussians(cloud, camera, config, use_sh=deg is not None, render_median_depth=extras)
    out.image.sum().backward()
    return out

  for _ in range(5):
    out = step()
  times = []
  for _ in range(steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = step()
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
  times.sort()
  ms = times[len(times) // 2]
  v = out.points.idx.shape[0]
  o2p, ranges = ts.map_to_tiles(out.points.gaussians2d.detach(), ts.rendering.ndc_depth(out.points.depths.detach(), camera.near_plane, camera.far_plane), size, config)
  r = ranges.view(-1, 2)
  per_tile = (r[:, 1] - r[:, 0])
  print(f"{name}\n    V={v} K={o2p.shape[0]} tiles={r.shape[0]} overlaps/tile mean {per_tile.float().mean():.0f} max {int(per_tile.max())}"
        f"  ->  {ms:.3f} ms/step, {n / ms / 1e3:.1f} M Gaussians/s")
  del cloud, out, o2p, ranges
  torch.cuda.empty_cache()
