"""A/B helper: wall time per forward+backward step of the bench workload for the library variant selected by
GS_BUILD_VARIANT (e.g. `GS_BUILD_VARIANT=_head python profiles/ab_step.py 200`).  Prints one line."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import taichi_splatting_b200 as ts
from taichi_splatting_b200.benchmarks import scenes

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda:0")
cam = scenes.benchmark_camera((2048, 2048))
cloud = scenes.random_3d_gaussians(1_000_000, cam, sh_degree=3, seed=0).to(dev).requires_grad_(True)
camera = cam.to(device=dev)
config = ts.RasterConfig(compute_visibility=True, compute_point_heuristic=True)


def step():
  out = ts.render_gaussians(cloud, camera, config, use_sh=True, render_median_depth=True)
  out.image.sum().backward()


for _ in range(10):
  step()
times = []
for _ in range(4):
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  torch.cuda.synchronize()
  a.record()
  for _ in range(steps // 4):
    step()
  b.record()
  torch.cuda.synchronize()
  times.append(a.elapsed_time(b) / (steps // 4))
switches = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("GS_") and k != "GS_BUILD_VARIANT")
print(f"variant '{os.environ.get('GS_BUILD_VARIANT', '')}' {switches}: {min(times):.4f} ms/step (quarters: {' '.join(f'{t:.4f}' for t in times)})")
