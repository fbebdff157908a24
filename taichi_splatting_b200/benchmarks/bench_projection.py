"""Projection micro-benchmark, the reference's CLI (taichi_splatting/benchmarks/bench_projection.py:18-88):
forward (project + cull + ordered compaction), then the backward for Gaussian parameters / extrinsics / intrinsics /
everything.

  python -m taichi_splatting_b200.benchmarks.bench_projection --n 2000000 --image_size 1024,768
"""
import argparse
import math

import torch

from ..data_types import RasterConfig
from ..perspective import CameraParams, project_to_image
from .scenes import benchmark_camera, random_3d_gaussians
from .util import benchmarked, size_arg


def random_view(image_size, seed: int) -> CameraParams:
  """A camera with a random pose, field of view and principal point (same ranges as the reference's test generator,
  tests/random_data.py:15-45)."""
  gen = torch.Generator().manual_seed(seed)
  w, h = image_size
  q = torch.nn.functional.normalize(torch.randn(4, generator=gen), dim=0)
  x, y, z, s = [float(v) for v in q]
  R = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - s * z), 2 * (x * z + s * y)],
                    [2 * (x * y + s * z), 1 - 2 * (x * x + z * z), 2 * (y * z - s * x)],
                    [2 * (x * z - s * y), 2 * (y * z + s * x), 1 - 2 * (x * x + y * y)]])
  T_world_camera = torch.eye(4)
  T_world_camera[:3, :3], T_world_camera[:3, 3] = R, torch.randn(3, generator=gen)
  fov = math.radians(float(torch.rand(1, generator=gen)) * 70 + 30)
  c = torch.tensor([w / 2, h / 2]) + torch.randn(2, generator=gen) * (w / 20)
  f = w / (2 * math.tan(fov / 2)), h / (2 * math.tan(fov / 2))
  return CameraParams(projection=torch.tensor([f[0], f[1], float(c[0]), float(c[1])]),
                      T_camera_world=torch.inverse(T_world_camera), near_plane=0.1, far_plane=100.0, image_size=(w, h))


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument("--profile", action="store_true")
  ap.add_argument("--image_size", type=str, default="1024,768")
  ap.add_argument("--device", type=str, default="cuda:0")
  ap.add_argument("--n", type=int, default=2000000)
  ap.add_argument("--seed", type=int, default=0)
  ap.add_argument("--iters", type=int, default=1000)
  ap.add_argument("--margin", type=float, default=0.5, help="controls random points (non visible) margin")
  ap.add_argument("--fixed_camera", action="store_true", help="the bench.py camera (identity pose, fov 60) instead of a random one")
  args = ap.parse_args(argv)
  size = size_arg(args.image_size)

  camera = benchmark_camera(size) if args.fixed_camera else random_view(size, args.seed)
  gaussians = random_3d_gaussians(args.n, camera, margin=args.margin, seed=args.seed).to(args.device)
  camera = camera.to(device=args.device)
  config = RasterConfig()
  with torch.no_grad():
    _, _, vis_idx = project_to_image(gaussians, camera, config)
    print(args)
    print(f"benchmarking {args.n} points ({vis_idx.shape[0]} visible) points")
    results = {"forward": benchmarked("forward", lambda: project_to_image(gaussians, camera, config), iters=args.iters,
                                      profile=args.profile)}

  def backward():
    for t in (*gaussians.shape_tensors(), camera.T_camera_world, camera.projection):
      t.grad = None
    points, depth, _ = project_to_image(gaussians, camera, config)
    (points.sum() + depth.sum()).backward()

  def grads(gauss: bool, extrinsics: bool, intrinsics: bool):
    for t in gaussians.shape_tensors():
      t.requires_grad_(gauss)
    camera.T_camera_world.requires_grad_(extrinsics)
    camera.projection.requires_grad_(intrinsics)

  for name, flags in (("backward (gaussians)", (True, False, False)), ("backward (extrinsics)", (False, True, False)),
                      ("backward (intrinsics)", (False, False, True)), ("backward (everything)", (True, True, True))):
    grads(*flags)
    results[name] = benchmarked(name, backward, iters=args.iters, profile=args.profile)
  return results


if __name__ == "__main__":
  main()
