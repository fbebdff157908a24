"""Does Morton-ordering the cloud (N4, misc/morton_sort.py) help the render path?  Times render_gaussians fwd+bwd at
cfg3 on the generator's (random) order and on the same cloud permuted into Morton order of the 3D positions."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import taichi_splatting_b200 as ts
from taichi_splatting_b200.benchmarks import scenes
from taichi_splatting_b200.misc import morton_sort

dev = torch.device("cuda:0")
size, n = (2048, 2048), 1_000_000
cam = scenes.benchmark_camera(size)
cloud = scenes.random_3d_gaussians(n, cam, sh_degree=3, seed=0).to(dev)
camera = cam.to(device=dev)
config = ts.RasterConfig(compute_visibility=True, compute_point_heuristic=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
order = morton_sort.argsort(cloud.position, 0.01).long()
b.record()
torch.cuda.synchronize()
print(f"morton argsort of {n} points: {a.elapsed_time(b):.3f} ms (first call)")
for name, c in (("generator order", cloud), ("Morton order", cloud.apply(lambda t: t[order].contiguous()))):
  c = c.requires_grad_(True)

  def step():
    for t in c.to_dict().values():
      t.grad = None
    out = ts.render_gaussians(c, camera, config, use_sh=True, render_median_depth=True)
    out.image.sum().backward()

  for _ in range(5):
    step()
  times = []
  for _ in range(30):
    a.record()
    step()
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
  print(f"{name}: {sorted(times)[15]:.3f} ms per fwd+bwd step")
