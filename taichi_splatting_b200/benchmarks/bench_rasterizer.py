"""Rasteriser micro-benchmark with the reference's CLI and protocol
(taichi_splatting/benchmarks/bench_rasterizer.py:20-108, benchmarks/util.py:23-37: 10 warm-ups, CUDA events, one sync).

  python -m taichi_splatting_b200.benchmarks.bench_rasterizer --n 1000000 --image_size 1024,768
"""
import argparse
from dataclasses import replace
from functools import partial

import torch

from ..data_types import RasterConfig
from ..mapper.tile_mapper import map_to_tiles
from ..misc.renderer2d import project_gaussians2d
from ..rasterizer import rasterize_with_tiles
from .scenes import random_2d_gaussians


def timed_benchmark(name, f, iters=100, warmup=10):
  for _ in range(warmup):
    f()
  start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  start.record()
  for _ in range(iters):
    f()
  end.record()
  torch.cuda.synchronize()
  elapsed = start.elapsed_time(end) / 1000.
  print(f"{name}  {iters} iterations in {elapsed:.3f}s at {iters / elapsed:.1f} iters/sec")
  return iters / elapsed


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument("--image_size", type=str, default="1024,768")
  ap.add_argument("--device", type=str, default="cuda:0")
  ap.add_argument("--n", type=int, default=1000000)
  ap.add_argument("--num_channels", type=int, default=3)
  ap.add_argument("--scale_factor", type=int, default=4)
  ap.add_argument("--tile_size", type=int, default=16)
  ap.add_argument("--seed", type=int, default=0)
  ap.add_argument("--iters", type=int, default=100)
  ap.add_argument("--antialias", action="store_true")
  ap.add_argument("--saturate_threshold", type=float, default=0.9999)
  ap.add_argument("--alpha_threshold", type=float, default=1 / 255)
  args = ap.parse_args(argv)
  size = tuple(map(int, args.image_size.split(",")))

  g = random_2d_gaussians(args.n, size, num_channels=args.num_channels, scale_factor=args.scale_factor,
                          alpha_range=(0.75, 1.0), depth_range=(0.1, 100.), seed=args.seed).to(args.device)
  config = RasterConfig(tile_size=args.tile_size, antialias=args.antialias, saturate_threshold=args.saturate_threshold,
                        alpha_threshold=args.alpha_threshold)
  g2d = project_gaussians2d(g)
  o2p, ranges = map_to_tiles(g2d, g.depths, image_size=size, config=config)
  per_tile = (ranges[:, :, 1] - ranges[:, :, 0]).float()
  print(f"scale_factor={args.scale_factor}, n={args.n}, tile_size={args.tile_size} "
        f"point_overlap={per_tile.sum().item() / args.n:.2f} tile_points={per_tile.mean().item():.2f}")

  def forward(cfg=config, gaussians2d=g2d, features=g.feature):
    return rasterize_with_tiles(gaussians2d, features, tile_overlap_ranges=ranges.view(-1, 2), overlap_to_point=o2p,
                                image_size=size, config=cfg)

  timed_benchmark("forward", forward, iters=args.iters * 4)
  timed_benchmark("forward_vis", partial(forward, replace(config, compute_visibility=True)), iters=args.iters * 4)

  def backward(cfg=config, grad_points=True, grad_features=True):
    p = g2d.detach().requires_grad_(grad_points)
    f = g.feature.detach().requires_grad_(grad_features)
    forward(cfg, p, f).image.sum().backward()

  timed_benchmark("backward (features)", partial(backward, grad_points=False), iters=args.iters)
  timed_benchmark("backward (gaussians)", partial(backward, grad_features=False), iters=args.iters)
  timed_benchmark("backward (all)", backward, iters=args.iters)
  timed_benchmark("backward (compute_point_heuristic)", partial(backward, replace(config, compute_point_heuristic=True)),
                  iters=args.iters)


if __name__ == "__main__":
  main()
