"""Work counters of the raster backward sweep (debug build: GS_BUILD_VARIANT=_count GS_BUILD_FLAGS=-DGS_COUNT).

Prints, for the bench workload, per visible Gaussian: warp iterations executed (hit list padded to the chunk),
hit-list entries (warp rectangle x splat pairs that survive the staging classification) and live (pixel, splat)
lanes -- i.e. how much of the sweep is useful work.
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import taichi_splatting_b200 as ts
from taichi_splatting_b200 import _lib
from taichi_splatting_b200.benchmarks import scenes

n = int(os.environ.get("GS_N", 1_000_000))
size = (int(os.environ.get("GS_W", 2048)), int(os.environ.get("GS_H", 2048)))
dev = torch.device("cuda:0")
cam = scenes.benchmark_camera(size)
cloud = scenes.random_3d_gaussians(n, cam, sh_degree=3, seed=0).to(dev).requires_grad_(True)
config = ts.RasterConfig(compute_visibility=True, compute_point_heuristic=True)
lib = ctypes.CDLL(str(_lib.LIB_PATH))
buf = (ctypes.c_ulonglong * 4)()
lib.gs_debug_counters(buf, 1)
out = ts.render_gaussians(cloud, cam.to(device=dev), config, use_sh=True, render_median_depth=True)
out.image.sum().backward()
torch.cuda.synchronize()
lib.gs_debug_counters(buf, 1)
v = out.points.idx.shape[0]
it, hits, live, batches = [int(x) for x in buf]
print(f"V={v} batches={batches} warp_iterations={it} ({it / v:.2f}/Gaussian) hit_entries={hits} ({hits / v:.2f}/Gaussian) "
      f"live_lanes={live} ({live / v:.1f}/Gaussian, {100 * live / (64 * max(it, 1)):.1f}% of the 64 pixels swept per iteration)")
