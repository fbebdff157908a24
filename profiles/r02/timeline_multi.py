"""GPU timeline (torch.profiler) of one view-parallel bench step on rank 0: kernels and NCCL collectives with their
stream, start, duration, to see what the exchange costs and what it overlaps.
  torchrun --nproc-per-node N profiles/r02/timeline_multi.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import taichi_splatting_b200 as ts
from taichi_splatting_b200 import parallel
from taichi_splatting_b200.benchmarks import scenes

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
size = (2048, 2048)
cloud = scenes.random_3d_gaussians(1_000_000, scenes.benchmark_camera(size), sh_degree=3, seed=0).to(dev).requires_grad_(True)
camera = scenes.benchmark_camera(size, yaw_deg=2.0 * rank).to(device=dev)
config = ts.RasterConfig(compute_visibility=True, compute_point_heuristic=True)


def step():
  for t in cloud.to_dict().values():
    t.grad = None
  out = parallel.render_view_parallel(cloud, camera, config, use_sh=True, render_median_depth=True, reduce_in_backward=True)
  out.image.sum().backward()


for _ in range(5):
  step()
dist.barrier()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
  for _ in range(3):
    step()
  torch.cuda.synchronize()
if rank == 0:
  evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
  evs.sort(key=lambda e: e.time_range.start)
  starts = [i for i, e in enumerate(evs) if "DeviceSelectSweep" in e.name]
  evs = evs[starts[1] - 1:starts[2] - 1]
  t0 = evs[0].time_range.start
  for e in evs:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    if d > 6:
      print(f"{s:9.1f} us  dur {d:8.1f}  {e.name[:90]}")
  print(f"step span {max(e.time_range.end for e in evs) - t0:.1f} us")
dist.destroy_process_group()
