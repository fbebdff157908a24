"""Tile-mapper micro-benchmark, the reference's CLI (taichi_splatting/benchmarks/bench_tilemapper.py:14-72): overlap
statistics of a random 2D cloud, then map_to_tiles (count, scan, key emit, sort, ranges) timed.

  python -m taichi_splatting_b200.benchmarks.bench_tilemapper --n 1000000 --scale_factor 2
"""
import argparse

from ..data_types import RasterConfig
from ..mapper.tile_mapper import map_to_tiles, map_to_tiles_full
from ..misc.renderer2d import project_gaussians2d
from .scenes import random_2d_gaussians
from .util import benchmarked, size_arg


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument("--profile", action="store_true")
  ap.add_argument("--image_size", type=str, default="1024,768")
  ap.add_argument("--device", type=str, default="cuda:0")
  ap.add_argument("--n", type=int, default=1000000)
  ap.add_argument("--scale_factor", type=float, default=2)
  ap.add_argument("--tile_size", type=int, default=16)
  ap.add_argument("--seed", type=int, default=0)
  ap.add_argument("--iters", type=int, default=1000)
  ap.add_argument("--depth16", action="store_true")
  ap.add_argument("--reference_sort", action="store_true",
                  help="also time the reference's own sequence: 64-bit (tile | depth) keys, one 48-bit radix sort")
  args = ap.parse_args(argv)
  size = size_arg(args.image_size)
  gaussians = random_2d_gaussians(args.n, size, scale_factor=args.scale_factor, alpha_range=(0.5, 1.0),
                                  depth_range=(0.1, 100.), seed=args.seed).to(args.device)
  config = RasterConfig(tile_size=args.tile_size)
  g2d = project_gaussians2d(gaussians)

  def run():
    return map_to_tiles(g2d, depth=gaussians.depths, image_size=size, config=config, use_depth16=args.depth16)

  _, tile_ranges = run()
  per_tile = tile_ranges[:, :, 1] - tile_ranges[:, :, 0]
  print(f"tile_mapper: scale_factor={args.scale_factor}, n={args.n}, tile_size={args.tile_size} "
        f"point_overlap={per_tile.sum().item() / args.n:.2f} tile_points={per_tile.float().mean().item():.2f}")
  results = {"tile_mapper": benchmarked("tile_mapper", run, iters=args.iters, profile=args.profile)}
  if args.reference_sort:
    results["tile_mapper (64-bit keys, one 48-bit sort)"] = benchmarked(
        "tile_mapper (64-bit keys, one 48-bit sort)",
        lambda: map_to_tiles_full(g2d, gaussians.depths, size, config, use_depth16=args.depth16, two_level=False),
        iters=args.iters, profile=args.profile)
  print("----------------------------------------------------------")
  return results


if __name__ == "__main__":
  main()
