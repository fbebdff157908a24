"""Whole render path on the CPU (oracle): the checker for render_gaussians and the CPU baseline.

TEST INFRASTRUCTURE ONLY.  Chains, in the order of the reference's renderer.py:22-108:
  torch_ops.project (autograd)  -> torch_ops.evaluate_sh_at | gather (autograd) -> torch_ops.ndc_depth
  -> cbind.map_to_tiles (C)     -> cbind.raster_forward (C)
and for the backward: cbind.raster_backward (C) -> torch autograd through SH / projection.
The projection + SH parts can run on the reference's own torch_lib (`use_reference=True`, build
container only) -- that is what `bench.py --impl reference` times when /root/reference is present.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from . import cbind, torch_ops


def render_forward_backward(g, camera, config, use_sh=False, grad_image=None, raster_dtype=np.float32,
                            use_reference=False, compute_grads=True, use_depth16=False):
  """g: namespace with position, log_scaling, rotation, alpha_logit, feature (CPU tensors).
  Returns a namespace with image, alpha, points, depths, indexes, overlap_to_point, tile_ranges,
  visibility, heuristic and (compute_grads) grads of every input under loss = sum(image * grad_image)."""
  leaves = {k: getattr(g, k).detach().clone().requires_grad_(compute_grads)
            for k in ("position", "log_scaling", "rotation", "alpha_logit", "feature")}
  Tcw = camera.T_camera_world.detach().clone().requires_grad_(compute_grads)
  proj = camera.projection.detach().clone().requires_grad_(compute_grads)

  if use_reference:
    from . import ref_loader
    ref_proj, ref_sh = ref_loader.load()
    project_fn, sh_fn = ref_proj.apply, ref_sh.evaluate_sh_at
  else:
    project_fn, sh_fn = torch_ops.project, torch_ops.evaluate_sh_at

  points, depths, indexes = project_fn(
      leaves["position"], leaves["log_scaling"], leaves["rotation"], leaves["alpha_logit"], Tcw, proj,
      camera.image_size, camera.depth_range, blur_cov=config.blur_cov, clamp_margin=config.clamp_margin,
      alpha_threshold=config.alpha_threshold)
  if use_sh:
    cam_pos = torch.inverse(Tcw)[0:3, 3]   # differentiable, as camera_params.camera_position is (perspective/params.py:78-80)
    features = sh_fn(leaves["feature"], leaves["position"].detach(), indexes, cam_pos)
  else:
    features = leaves["feature"][indexes]

  ndc = torch_ops.ndc_depth(depths.detach(), camera.near_plane, camera.far_plane)
  o2p, ranges = cbind.map_to_tiles(points.detach().float(), ndc.float(), camera.image_size, config,
                                   use_depth16=use_depth16)
  image, alpha, vis = cbind.raster_forward(points, features, ranges, o2p, camera.image_size, config,
                                           dtype=raster_dtype)
  out = SimpleNamespace(image=image, alpha=alpha, visibility=vis, points=points.detach(), depths=depths.detach(),
                        indexes=indexes, features=features.detach(), overlap_to_point=o2p, tile_ranges=ranges,
                        ndc=ndc, heuristic=None, grads=None)
  if not compute_grads:
    return out

  gi = np.ones_like(image) if grad_image is None else np.asarray(grad_image, dtype=image.dtype)
  gp, gf, heur = cbind.raster_backward(points, features, ranges, o2p, image, gi, camera.image_size, config,
                                       dtype=raster_dtype)
  out.heuristic = heur
  out.grad_points, out.grad_features = gp, gf
  torch.autograd.backward([points, features],
                          [torch.from_numpy(gp).to(points.dtype), torch.from_numpy(gf).to(features.dtype)])
  out.grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
  out.grads["T_camera_world"] = Tcw.grad if Tcw.grad is not None else torch.zeros_like(Tcw)
  out.grads["projection"] = proj.grad if proj.grad is not None else torch.zeros_like(proj)
  return out
