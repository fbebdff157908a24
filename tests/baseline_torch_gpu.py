"""B-torch-gpu: the reference's OWN pure-torch projection + spherical harmonics (torch_lib,
restated expression for expression in oracle/torch_ops.py and pinned to it by tests/golden) timed ON THE B200, forward +
backward, next to this package's kernels for the same stages on the same inputs.  It is the one GPU timing of reference
code for rows R1 / R1b / R2 that can be taken in this image (the Taichi kernels cannot run).

Not a pytest module (lives under tests/ because only tests may import oracle/):
    python tests/baseline_torch_gpu.py [iters]  > profiles/r02/r02_torch_gpu_baseline.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import taichi_splatting_b200 as ts
from oracle import torch_ops
from taichi_splatting_b200.benchmarks import scenes

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")


def timed(f, n=iters, warm=3):
  for _ in range(warm):
    f()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n):
    f()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / n


for name, n, size, deg in (("cfg2: 100 k Gaussians, 1024x1024, SH deg 0", 100_000, (1024, 1024), 0),
                           ("cfg3: 1 M Gaussians, 2048x2048, SH deg 3", 1_000_000, (2048, 2048), 3)):
  cam = scenes.benchmark_camera(size)
  cloud = scenes.random_3d_gaussians(n, cam, sh_degree=deg, seed=0).to(dev)
  geom = [t.detach().clone().requires_grad_(True) for t in cloud.shape_tensors()]
  sh = cloud.feature.detach().clone().requires_grad_(True)
  Tcw, proj = cam.T_camera_world.to(dev), cam.projection.to(dev)
  cam_pos = torch.inverse(Tcw)[0:3, 3]
  depth_range = (cam.near_plane, cam.far_plane)

  def ref_step():
    for t in (*geom, sh):
      t.grad = None
    pts, depth, idx = torch_ops.project(*geom, Tcw, proj, size, depth_range, blur_cov=0.3)
    feats = torch_ops.evaluate_sh_at(sh, geom[0].detach(), idx, cam_pos)
    (pts.sum() + depth.sum() + feats.sum()).backward()
    return idx

  def our_step():
    for t in (*geom, sh):
      t.grad = None
    pts, depth, idx = ts.perspective.apply(*geom, Tcw, proj, size, depth_range, blur_cov=0.3)
    feats = ts.evaluate_sh_at(sh, geom[0].detach(), idx, cam_pos, unique_indexes=True)
    (pts.sum() + depth.sum() + feats.sum()).backward()
    return idx

  v_ref, v_our = ref_step().shape[0], our_step().shape[0]
  assert v_ref == v_our, (v_ref, v_our)
  t_ref, t_our = timed(ref_step), timed(our_step)
  print(f"{name} (V = {v_our}): projection + SH, forward + backward incl. the loss kernels")
  print(f"    reference torch_lib arithmetic on the B200 (eager torch ops): {t_ref:8.3f} ms  = {n / t_ref / 1e3:8.1f} M Gaussians/s")
  print(f"    this package (gs_project_* + gs_sh_* kernels)              : {t_our:8.3f} ms  = {n / t_our / 1e3:8.1f} M Gaussians/s"
        f"   ({t_ref / t_our:.1f}x)")
