"""One render_gaussians forward+backward step of the bench workload, for ncu captures.

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python profiles/profile_step.py
  ncu --set full --clock-control none --import-source on -k regex:raster_ -o gpurun_out/raster \
      python profiles/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import taichi_splatting_b200 as ts
from taichi_splatting_b200.benchmarks import scenes

n = int(os.environ.get("GS_N", 1_000_000))
size = (int(os.environ.get("GS_W", 2048)), int(os.environ.get("GS_H", 2048)))
steps = int(os.environ.get("GS_STEPS", 1))
dev = torch.device("cuda:0")
cam = scenes.benchmark_camera(size)
cloud = scenes.random_3d_gaussians(n, cam, sh_degree=3, seed=0).to(dev).requires_grad_(True)
camera = cam.to(device=dev)
config = ts.RasterConfig(compute_visibility=True, compute_point_heuristic=True)
for _ in range(steps):
  for t in cloud.to_dict().values():   # as bench.py: no gradient accumulation kernels in the launch list
    t.grad = None
  out = ts.render_gaussians(cloud, camera, config, use_sh=True, render_median_depth=True)
  out.image.sum().backward()
torch.cuda.synchronize()
print("V", out.points.idx.shape[0])
