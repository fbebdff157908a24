"""GPU timeline of one bench step from torch.profiler (kernel start/duration/idle gap), to find GPU idle time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import taichi_splatting_b200 as ts
from taichi_splatting_b200.benchmarks import scenes

dev = torch.device("cuda:0")
n = int(os.environ.get("GS_N", 1_000_000))
size = (int(os.environ.get("GS_W", 2048)), int(os.environ.get("GS_H", 2048)))
deg = int(os.environ.get("GS_DEG", 3))          # -1: plain features
extras = os.environ.get("GS_EXTRAS", "1") != "0"
cam = scenes.benchmark_camera(size)
cloud = scenes.random_3d_gaussians(n, cam, sh_degree=deg if deg >= 0 else None, seed=0).to(dev).requires_grad_(True)
camera = cam.to(device=dev)
config = ts.RasterConfig(compute_visibility=extras, compute_point_heuristic=extras)


def step():
  for t in cloud.to_dict().values():
    t.grad = None
  out = ts.render_gaussians(cloud, camera, config, use_sh=deg >= 0, render_median_depth=extras)
  out.image.sum().backward()


for _ in range(5):
  step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
  step()
  step()
  torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# keep the second step only
mid = [i for i, e in enumerate(evs) if "DeviceCompactInitKernel" in e.name]   # first kernel of gs_project_compact
evs = evs[mid[1]:]
t0 = evs[0].time_range.start
prev_end = t0
busy = 0.0
for e in evs:
  s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
  gap = e.time_range.start - prev_end
  busy += d
  if d > float(os.environ.get('GS_MIN_US', 8)) or gap > 8:
    print(f"{s:9.1f} us  dur {d:8.1f}  idle-before {gap:7.1f}  {e.name[:70]}")
  prev_end = max(prev_end, e.time_range.end)
print(f"step span {prev_end - t0:.1f} us, busy {busy:.1f} us, idle {prev_end - t0 - busy:.1f} us")
