"""Per-stage CUDA-event timing of the bench workload (median over steps).  Usage: python profiles/time_stages.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import taichi_splatting_b200 as ts
from taichi_splatting_b200 import _lib
from taichi_splatting_b200.benchmarks import scenes

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = int(os.environ.get("GS_N", 1_000_000))
size = (int(os.environ.get("GS_W", 2048)), int(os.environ.get("GS_H", 2048)))
dev = torch.device("cuda:0")
cam = scenes.benchmark_camera(size)
cloud = scenes.random_3d_gaussians(n, cam, sh_degree=3, seed=0).to(dev).requires_grad_(True)
camera = cam.to(device=dev)
config = ts.RasterConfig(compute_visibility=True, compute_point_heuristic=True)


def step():
  out = ts.render_gaussians(cloud, camera, config, use_sh=True, render_median_depth=True)
  out.image.sum().backward()


for _ in range(5):
  step()
prof = _lib.Profiler()
_lib.profiler = prof
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
  step()
b.record()
torch.cuda.synchronize()
_lib.profiler = None
tot = 0.0
for k, v in prof.stage_ms().items():
  v = sorted(v)
  med = v[len(v) // 2]
  tot += med
  if med > 0.004:
    print(f"{k:32s} median {med:8.4f} ms   min {v[0]:8.4f}")
print(f"sum of stage medians {tot:.4f} ms; wall per step {a.elapsed_time(b) / steps:.4f} ms")

# ---- timeline of one step: where the GPU idles between our launches ----
prof = _lib.Profiler()
_lib.profiler = prof
t0 = torch.cuda.Event(enable_timing=True)
t1 = torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
t0.record()
step()
t1.record()
torch.cuda.synchronize()
_lib.profiler = None
prev_end = 0.0
print("timeline (ms from step start): name start end gap_before")
for name, a_, b_ in prof.records:
  s, e = t0.elapsed_time(a_), t0.elapsed_time(b_)
  if e - s > 0.004 or s - prev_end > 0.02:
    print(f"  {name:28s} {s:8.3f} {e:8.3f}  gap {s - prev_end:7.3f}")
  prev_end = e
print(f"  step end {t0.elapsed_time(t1):8.3f}  gap {t0.elapsed_time(t1) - prev_end:7.3f}")
