"""Generates tests/golden/*.npz from the REAL reference torch_lib (run in the build container).

  python tests/golden/make_golden.py

Inputs come from oracle/random_data.py (restated reference generators, seeds below); outputs and
autograd gradients come from /root/reference/taichi_splatting/torch_lib (projection.apply,
spherical_harmonics.evaluate_sh_at), loaded through oracle/ref_loader.py.  Loss for gradients is
the one the reference's own tests use (tests/util.py:10-33): sum of mean() of every float output.
The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import random_data, ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def projection_case(proj, seed, max_points, dtype, scale_factor=0.1, margin=0.5, blur_cov=0.0):
  # tests/test_projection.py:22-33 random_inputs
  torch.manual_seed(seed)
  camera = random_data.random_camera()
  n = torch.randint(size=(1,), low=1, high=max_points).item()
  g = random_data.random_3d_gaussians(n=n, camera=camera, margin=margin, scale_factor=scale_factor)
  ins = [g.position, g.log_scaling, g.rotation, g.alpha_logit, camera.T_camera_world, camera.projection]
  ins = [x.to(dtype).detach().clone().requires_grad_(True) for x in ins]
  points, depth, idx = proj.apply(*ins, camera.image_size, camera.depth_range, blur_cov=blur_cov)
  (points.mean() + depth.mean()).backward()
  names = ["position", "log_scaling", "rotation", "alpha_logit", "T_camera_world", "projection"]
  out = {f"in_{k}": v.detach().numpy() for k, v in zip(names, ins)}
  out.update({f"grad_{k}": v.grad.numpy() for k, v in zip(names, ins)})
  out.update(points=points.detach().numpy(), depth=depth.detach().numpy(), indexes=idx.numpy(),
             image_size=np.array(camera.image_size), depth_range=np.array(camera.depth_range),
             blur_cov=np.array(blur_cov))
  return out


def sh_case(sh, seed, dtype):
  # tests/test_spherical_harmonics.py:15-31 random_inputs
  torch.random.manual_seed(seed)
  dimension = torch.randint(1, 4, (1,)).item()
  degree = torch.randint(1, 4, (1,)).item()
  n = torch.randint(1, 102, (1,)).item()
  params = torch.rand(n, dimension, (degree + 1)**2, dtype=dtype)
  points = torch.randn(n, 3, dtype=dtype)
  camera_pos = torch.randn(3, dtype=dtype)
  indexes = torch.randint(0, n, (n // 2,))
  ins = [params.requires_grad_(True), points.requires_grad_(True), camera_pos.requires_grad_(True)]
  out = sh.evaluate_sh_at(ins[0], ins[1], indexes, ins[2])
  out.mean().backward()
  return dict(in_params=params.detach().numpy(), in_points=points.detach().numpy(),
              in_camera_pos=camera_pos.detach().numpy(), indexes=indexes.numpy(),
              out=out.detach().numpy(), grad_params=ins[0].grad.numpy(),
              grad_points=ins[1].grad.numpy(), grad_camera_pos=ins[2].grad.numpy())


def main():
  assert ref_loader.available(), "reference tree not present"
  proj, sh = ref_loader.load()
  cases = {}
  for seed in range(8):  # float64, as the reference's test_projection does
    for k, v in projection_case(proj, seed, 400, torch.float64).items():
      cases[f"p64_{seed}_{k}"] = v
  for seed in range(4):  # float32 + the renderer's default blur
    for k, v in projection_case(proj, 100 + seed, 400, torch.float32, scale_factor=1.0, margin=0.2, blur_cov=0.3).items():
      cases[f"p32_{seed}_{k}"] = v
  np.savez_compressed(os.path.join(HERE, "projection.npz"), **cases)
  cases = {}
  for seed in range(12):
    for k, v in sh_case(sh, seed, torch.float32 if seed % 2 else torch.float64).items():
      cases[f"sh_{seed}_{k}"] = v
  np.savez_compressed(os.path.join(HERE, "spherical_harmonics.npz"), **cases)
  for f in ("projection.npz", "spherical_harmonics.npz"):
    print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
  main()
