"""2D image fitting with Gaussians: the reference's only end-to-end training entry, reduced to the render path.

Reference: taichi_splatting/examples/fit_image_gaussians.py:89-147,234-358 (BASELINE config 1: 256x256, n=2000).
Same forward / loss / backward through `rasterize` as the reference, and by default the reference's optimiser setup
(:267-281): VisibilityAwareLaProp over the visible points with a `local_vector` position group stepped in each
Gaussian's own basis, through `taichi_splatting_b200.optim`.  `--opt adam` uses torch.optim.Adam instead.
`--target N` turns on the reference's densification (:190-230, 299-340): training runs in epochs; the raster backward's
per-point heuristics (prune cost, split score) are accumulated over an epoch, and between epochs the least useful
points are pruned and the most pulled-on ones split until the cloud has N points (`misc/densify.py`), the sparse
optimiser's per-point state following its rows.

  python -m taichi_splatting_b200.examples.fit_image_gaussians [--image path.png] [--n 2000] [--iters 200] [--opt laprop]
      [--target 4000 --epoch 50 --prune_rate 0.04]
"""
import argparse
import math

import torch

from ..benchmarks.scenes import random_2d_gaussians
from ..data_types import RasterConfig
from ..misc.densify import split_prune
from ..misc.renderer2d import point_basis, project_gaussians2d
from ..optim import SparseAdam, VisibilityAwareLaProp
from ..rasterizer import rasterize


def synthetic_image(w, h, device, detail=1.0):
  """Smooth colour waves; `detail` scales their spatial frequency (more detail needs more Gaussians)."""
  ys, xs = torch.meshgrid(torch.linspace(0, 1, h, device=device), torch.linspace(0, 1, w, device=device), indexing="ij")
  return torch.stack([0.5 + 0.5 * torch.sin(6.28 * detail * xs), 0.5 + 0.5 * torch.sin(3.14 * (2 * detail - 1) * ys) if detail != 1.0 else ys,
                      0.5 + 0.5 * torch.cos(9.4 * detail * detail * xs * ys)], dim=-1).contiguous()


def psnr(a, b):
  return 10 * math.log10(1 / torch.nn.functional.mse_loss(a, b).item())


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument("--image", type=str, default=None, help="optional image file (needs cv2); default: synthetic")
  ap.add_argument("--n", type=int, default=2000)
  ap.add_argument("--size", type=str, default="256,256")
  ap.add_argument("--detail", type=float, default=1.0, help="spatial frequency of the synthetic target")
  ap.add_argument("--iters", type=int, default=200)
  ap.add_argument("--lr", type=float, default=0.02)
  ap.add_argument("--tile_size", type=int, default=16)
  ap.add_argument("--opt", type=str, default="laprop", choices=["laprop", "sparse_adam", "adam"],
                  help="laprop: VisibilityAwareLaProp (the reference's choice); sparse_adam: SparseAdam; adam: torch Adam")
  ap.add_argument("--device", type=str, default="cuda:0")
  ap.add_argument("--target", type=int, default=None, help="densify: split / prune between epochs until the cloud has this many points")
  ap.add_argument("--epoch", type=int, default=50, help="iterations per densification epoch")
  ap.add_argument("--prune_rate", type=float, default=0.04, help="fraction of points pruned per epoch (decays to 0 over training)")
  args = ap.parse_args(argv)
  device = torch.device(args.device)

  if args.image is not None:
    import cv2
    img = cv2.cvtColor(cv2.imread(args.image), cv2.COLOR_BGR2RGB)
    ref_image = (torch.from_numpy(img).to(torch.float32) / 255).to(device)
  else:
    w, h = map(int, args.size.split(","))
    ref_image = synthetic_image(w, h, device, args.detail)
  h, w = ref_image.shape[:2]

  g = random_2d_gaussians(args.n, (w, h), alpha_range=(0.5, 1.0), scale_factor=0.5, seed=0).to(device)
  params = [g.position, g.log_scaling, g.rotation, g.alpha_logit, g.feature]
  for p in params:
    p.requires_grad_(True)
  if args.opt == "adam":
    opt = torch.optim.Adam([{"params": [g.position], "lr": args.lr * 10}, {"params": params[1:], "lr": args.lr}])
  else:   # parameter groups of the reference (:267-274)
    groups = [dict(params=[g.position], name="position", lr=0.5, type="local_vector"),
              dict(params=[g.log_scaling], name="log_scaling", lr=0.1, type="scalar"),
              dict(params=[g.rotation], name="rotation", lr=1.0, type="scalar"),
              dict(params=[g.alpha_logit], name="alpha_logit", lr=0.1, type="scalar"),
              dict(params=[g.feature], name="feature", lr=0.025, type="vector")]
    if args.opt == "laprop":
      opt = VisibilityAwareLaProp(groups, vis_smooth=0.1, vis_beta=0.8, betas=(0.9, 0.9), eps=1e-16, bias_correction=True)
    else:
      opt = SparseAdam(groups, betas=(0.9, 0.95), eps=1e-16, bias_correction=True)
  densify = args.target is not None
  assert not (densify and args.opt == "adam"), "densification carries per-point optimiser state: use --opt laprop | sparse_adam"
  config = RasterConfig(tile_size=args.tile_size, compute_visibility=True, compute_point_heuristic=densify)
  heuristics = torch.zeros((g.batch_size[0], 2), device=device)

  for it in range(args.iters):
    if densify and it > 0 and it % args.epoch == 0 and it + args.epoch <= args.iters:
      # between epochs: prune by accumulated prune cost, split by accumulated split score (reference :299-340)
      n_before = g.batch_size[0]
      g, info = split_prune(g, t=it / args.iters, target=args.target, prune_rate=args.prune_rate,
                            heuristics=(heuristics[:, 0], heuristics[:, 1]), optimizer=opt)
      heuristics = torch.zeros((g.batch_size[0], 2), device=device)
      print(f"iter {it:5d}  densify: {n_before} -> {g.batch_size[0]} points (split {info['split']}, pruned {info['prune']})")
    opt.zero_grad(set_to_none=True)
    raster = rasterize(project_gaussians2d(g), torch.clamp(g.depths, 0, 1), g.feature, (w, h), config)
    loss = torch.nn.functional.mse_loss(raster.image, ref_image)
    loss.backward()
    if densify:
      heuristics += raster.point_heuristic
    if args.opt == "adam":
      opt.step()
    else:   # step the visible points only (:125-136)
      visible = (raster.visibility > 1e-8).nonzero().squeeze(1)
      with torch.no_grad():
        basis = point_basis(g[visible])
      if args.opt == "laprop":
        opt.step(indexes=visible, visibility=raster.visibility[visible], basis=basis)
      else:
        opt.step(indexes=visible, basis=basis)
    if it % 50 == 0 or it == args.iters - 1:
      visible = int((raster.visibility > 0).sum())
      print(f"iter {it:5d}  psnr {psnr(raster.image.detach(), ref_image):6.2f} dB  visible {visible}/{g.batch_size[0]}")
  return psnr(raster.image.detach(), ref_image)


if __name__ == "__main__":
  main()
