"""Numerics of the tile-centred affine form used by the tuned raster kernels (CPU, numpy):
(tx, ty) = X (ux, wx) + Y (uy, wy) + (tx0, ty0) in fp32 against the reference form (d . axis / sigma) in fp64, over the
bench workload's splat distribution.  Prints the error distribution of alpha = alpha_point * exp(-0.5 (tx^2 + ty^2)) on
the contributing pixels (alpha > 1/255), relative to alpha and relative to the image scale (absolute)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import torch_ops
from oracle.cbind import OracleConfig
from taichi_splatting_b200.benchmarks import scenes

size, n = (2048, 2048), 1_000_000
cam = scenes.benchmark_camera(size)
g = scenes.random_3d_gaussians(n, cam, sh_degree=None, seed=0)
cfg = OracleConfig()
with torch.no_grad():
  pts, _, _ = torch_ops.project(g.position, g.log_scaling, g.rotation, g.alpha_logit, cam.T_camera_world, cam.projection,
                                size, cam.depth_range, blur_cov=cfg.blur_cov, clamp_margin=cfg.clamp_margin,
                                alpha_threshold=cfg.alpha_threshold)
pts = pts.numpy().astype(np.float32)
rng = np.random.default_rng(0)
f32, k, thr = np.float32, np.float32(0.84932180028801904), 1.0 / 255.0
rel, absolute = [], []
for i in rng.choice(len(pts), 3000, replace=False):
  mx, my, ax, ay, sx, sy, al = pts[i]
  r = float(np.sqrt(2 * np.log(max(al / thr, 1.0 + 1e-6))) * max(sx, sy)) + 1
  for tx in range(int(max(0, (mx - r) // 16)), int(min(127, (mx + r) // 16)) + 1):
    for ty in range(int(max(0, (my - r) // 16)), int(min(127, (my + r) // 16)) + 1):
      # fp64 reference form (forward.py / generic.py:310-317)
      X, Y = np.meshgrid(np.arange(16) + tx * 16 + 0.5, np.arange(16) + ty * 16 + 0.5)
      dx, dy = X - float(mx), Y - float(my)
      t1 = (dx * float(ax) + dy * float(ay)) / float(sx)
      t2 = (dy * float(ax) - dx * float(ay)) / float(sy)
      a_ref = float(al) * np.exp(-0.5 * (t1 * t1 + t2 * t2))
      # fp32 tile-centred affine form (raster_digest.cu + stage_splat + the sweep)
      isx, isy = f32(1) / sx, f32(1) / sy
      ux, uy, wx, wy = ax * isx * k, ay * isx * k, -ay * isy * k, ax * isy * k
      ddx, ddy = mx - f32(tx * 16 + 8), my - f32(ty * 16 + 8)
      tx0, ty0 = -(ux * ddx + uy * ddy), -(wx * ddx + wy * ddy)
      lx, ly = np.meshgrid((np.arange(16) - 7.5).astype(f32), (np.arange(16) - 7.5).astype(f32))
      tX = lx * ux + (ly * uy + tx0)
      tY = lx * wx + (ly * wy + ty0)
      a32 = al * np.exp2(-(tX * tX + tY * tY)).astype(f32)
      live = a_ref > thr
      if live.any():
        rel.append(np.abs(a32[live].astype(np.float64) - a_ref[live]) / a_ref[live])
        absolute.append(np.abs(a32[live].astype(np.float64) - a_ref[live]))
rel, absolute = np.concatenate(rel), np.concatenate(absolute)
print(f"{rel.size} contributing (pixel, splat) pairs of 3000 Gaussians of the bench cloud")
for name, e in (("relative error of alpha", rel), ("absolute error of alpha", absolute)):
  print(f"  {name}: median {np.median(e):.2e}  99 % {np.quantile(e, 0.99):.2e}  99.99 % {np.quantile(e, 0.9999):.2e}  max {e.max():.2e}")
