import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import taichi_splatting_b200 as ts
from taichi_splatting_b200.benchmarks import scenes
dev = torch.device("cuda:0")
size = (160, 96)
cam = scenes.benchmark_camera(size)
cloud = scenes.random_3d_gaussians(2000, cam, sh_degree=1, seed=0).to(dev)
far = cloud.replace(position=cloud.position - torch.tensor([0., 0., 1e4], device=dev)).requires_grad_(True)
cfg = ts.RasterConfig(compute_visibility=True)
out = ts.render_gaussians(far, cam.to(device=dev), cfg, use_sh=True)   # first frame of the process: V = K = 0
out.image.sum().backward()
assert out.points.idx.numel() == 0 and float(out.image.detach().abs().max()) == 0.0
near = cloud.requires_grad_(True)
out = ts.render_gaussians(near, cam.to(device=dev), cfg, use_sh=True)
out.image.sum().backward()
torch.cuda.synchronize()
print("ok", out.points.idx.numel(), float(out.image.detach().mean()))
