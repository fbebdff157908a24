"""Spherical-harmonics micro-benchmark, the reference's CLI (taichi_splatting/benchmarks/bench_sh.py:14-65): forward,
backward for the coefficients, backward for coefficients + positions + camera position.

  python -m taichi_splatting_b200.benchmarks.bench_sh --n 1000000 --degree 3
"""
import argparse

import torch

from ..spherical_harmonics import evaluate_sh_at
from .util import benchmarked, size_arg


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument("--profile", action="store_true")
  ap.add_argument("--image_size", type=str, default="1024,768")   # accepted and unused, as in the reference
  ap.add_argument("--device", type=str, default="cuda:0")
  ap.add_argument("--n", type=int, default=1000000)
  ap.add_argument("--seed", type=int, default=0)
  ap.add_argument("--iters", type=int, default=200)
  ap.add_argument("--degree", type=int, default=3)
  args = ap.parse_args(argv)
  size_arg(args.image_size)
  gen = torch.Generator().manual_seed(args.seed)
  sh = torch.randn(args.n, 3, (args.degree + 1)**2, generator=gen).to(args.device)
  points = torch.randn(args.n, 3, generator=gen).to(args.device)
  indexes = torch.arange(args.n, device=args.device)
  camera_pos = torch.zeros(3, device=args.device)
  print(args)
  results = {}
  with torch.no_grad():
    results["forward"] = benchmarked("forward", lambda: evaluate_sh_at(sh, points, indexes, camera_pos), iters=args.iters,
                                     profile=args.profile)

  def backward():
    for t in (sh, points, camera_pos):
      t.grad = None
    evaluate_sh_at(sh, points, indexes, camera_pos).sum().backward()

  sh.requires_grad_(True)
  results["backward (sh_features)"] = benchmarked("backward (sh_features)", backward, iters=args.iters, profile=args.profile)
  points.requires_grad_(True)
  camera_pos.requires_grad_(True)
  results["backward (all)"] = benchmarked("backward (all)", backward, iters=args.iters, profile=args.profile)
  return results


if __name__ == "__main__":
  main()
