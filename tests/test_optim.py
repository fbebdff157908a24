"""N3 -- sparse / visibility-aware optimisers.

Golden vectors (tests/golden/optim.npz) come from the REAL reference host code (optim/fractional.py,
optim/visibility_aware.py) running over the restated Taichi kernels (tests/golden/make_golden_optim.py).
CPU: the oracle's restated host logic against them (and, in the build container, a live re-run of the reference).
GPU: the CUDA optimisers (gs_optim_step_f32, gs_optim_update_visibility_f32) against them."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import optim_ops

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_optim", os.path.join(HERE, "golden", "make_golden_optim.py"))
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)
GOLDEN = np.load(os.path.join(HERE, "golden", "optim.npz"))


def _check(name, recs, tol):
  for s, rec in enumerate(recs):
    for k, v in rec.items():
      ref = GOLDEN[f"{name}/{s}/{k}"]
      err = np.abs(v.astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30)
      assert err < tol, (name, s, k, err)
    assert {f"{name}/{s}/{k}" for k in rec} == {k for k in GOLDEN.files if k.startswith(f"{name}/{s}/")}


class _OracleOptimiser:
  """The oracle's restated host logic behind the optimiser interface make_golden_optim.run drives."""

  def __init__(self, groups, algorithm, visibility=False, vis_beta=0.5, vis_smooth=0.01, lr=0.01):
    self.groups, self.algorithm, self.visibility = groups, algorithm, visibility
    self.vis_beta, self.vis_smooth = vis_beta, vis_smooth
    self.state = {g["params"][0]: {} for g in groups}
    for g in groups:
      g.setdefault("betas", (0.9, 0.999)), g.setdefault("eps", 1e-16), g.setdefault("bias_correction", True)

  @torch.no_grad()
  def step(self, indexes, weight=None, basis=None):
    first = self.state[self.groups[0]["params"][0]]
    n = self.groups[0]["params"][0].shape[0]
    total = first.setdefault("total_weight", torch.zeros(n))
    scale = None
    if weight is None:
      weight = torch.ones(indexes.shape[0])
    if self.visibility:
      running = first.setdefault("running_vis", torch.zeros(n))
      scale = weight
      weight = optim_ops.update_visibility(running, scale, indexes, self.vis_beta)
    total[indexes] += weight
    for g in self.groups:
      p = g["params"][0]
      param, grad = p.view(p.shape[0], -1), p.grad.view(p.shape[0], -1)
      if scale is not None:
        scaled = torch.zeros_like(grad)
        scaled[indexes] = grad[indexes] / (scale.unsqueeze(1) + self.vis_smooth)
        grad = scaled
      optim_ops.group_step(g, self.state[p], param, grad, indexes, weight, total, self.algorithm, basis)


@pytest.mark.parametrize("name", mk.OPTIMISERS)
def test_oracle_host_logic_matches_reference(name):
  init, mask_lr, point_lr, steps = mk.scenario()
  alg = optim_ops.LAPROP if "LaProp" in name else optim_ops.ADAM
  recs = mk.run(lambda groups: _OracleOptimiser(groups, alg, visibility=name.startswith("Visibility")),
                mk.kind_of(name), init, mask_lr, point_lr, steps)
  _check(name, recs, 1e-6)


@pytest.mark.skipif(not optim_ops.reference_available(), reason="reference tree not present")
def test_golden_reproducible_from_reference():
  fr, va = optim_ops.load_reference_optim()
  init, mask_lr, point_lr, steps = mk.scenario()
  for name in ("FractionalLaProp", "VisibilityAwareAdam"):
    cls = getattr(va if name.startswith("Visibility") else fr, name)
    _check(name, mk.run(lambda groups: cls(groups, lr=0.01), mk.kind_of(name), init, mask_lr, point_lr, steps), 1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("name", mk.OPTIMISERS)
def test_cuda_optimisers_match_reference(name):
  from taichi_splatting_b200 import optim
  init, mask_lr, point_lr, steps = mk.scenario()
  cls = getattr(optim, name)
  recs = mk.run(lambda groups: cls(groups, lr=0.01), mk.kind_of(name), init, mask_lr, point_lr, steps, device="cuda:0")
  _check(name, recs, 2e-5)


@pytest.mark.gpu
def test_cuda_optimiser_after_render():
  """The intended use: render -> backward -> visibility-aware step on the visible set (fit_image_gaussians.py:120-147)."""
  import taichi_splatting_b200 as ts
  from oracle import random_data
  from taichi_splatting_b200 import optim
  torch.manual_seed(1)
  size = (160, 96)
  cam = random_data.fixed_camera(size)
  g = random_data.random_3d_gaussians(3000, cam, scale_factor=1.5, margin=0.2)
  dev = "cuda:0"
  params = {k: v.to(dev).requires_grad_(True) for k, v in vars(g).items()}
  kinds = {"position": "vector", "log_scaling": "scalar", "rotation": "scalar", "alpha_logit": "scalar", "feature": "scalar"}
  opt = optim.VisibilityAwareLaProp([dict(params=[p], name=k, type=kinds[k], lr=1e-3) for k, p in params.items()], lr=1e-3)
  camera = ts.perspective.CameraParams(projection=cam.projection.to(dev), T_camera_world=cam.T_camera_world.to(dev),
                                       near_plane=cam.near_plane, far_plane=cam.far_plane, image_size=size)
  target = torch.rand((size[1], size[0], 3), device=dev)
  cfg = ts.RasterConfig(compute_visibility=True)
  losses = []
  for _ in range(12):
    for p in params.values():
      p.grad = None
    out = ts.render_gaussians(ts.Gaussians3D(**params, batch_size=(3000,)), camera, cfg)
    loss = ((out.image - target) ** 2).mean()
    loss.backward()
    vis = out.points.visibility
    keep = vis > 1e-8
    opt.step(out.points.idx[keep], vis[keep])
    losses.append(float(loss.detach()))
  assert losses[-1] < losses[0], losses
  assert all(torch.isfinite(p).all() for p in params.values())
