"""Renderer facade (reference: taichi_splatting/renderer.py:22-121): project -> SH | gather -> map_to_tiles
-> rasterize (-> median-depth raster), each stage one of this package's operators."""
import torch
from beartype import beartype

from .data_types import Gaussians3D, RasterConfig
from .mapper.tile_mapper import map_to_tiles
from .perspective import CameraParams
from .perspective.projection import apply_with_ndc, camera_position
from .rasterizer.function import rasterize_with_tiles, rasterize_with_tiles_and_median
from .rendering import RenderedPoints, Rendering, ndc_depth
from .spherical_harmonics import evaluate_sh_at


@beartype
def render_gaussians(gaussians: Gaussians3D, camera_params: CameraParams, config: RasterConfig = RasterConfig(),
                     use_sh: bool = False, render_depth: bool = False, use_depth16: bool = False,
                     render_median_depth: bool = False) -> Rendering:
  """Complete renderer for 3D Gaussians; same parameters and result type as the reference (:22-59).
  `render_depth` is accepted and unused, as in the reference."""
  gaussians2d, depths, indexes, ndc = apply_with_ndc(
      *gaussians.shape_tensors(), camera_params.T_camera_world, camera_params.projection,
      camera_params.image_size, camera_params.depth_range, config.blur_cov, config.clamp_margin,
      config.alpha_threshold)

  if use_sh:
    features = evaluate_sh_at(gaussians.feature, gaussians.position.detach(), indexes,
                              camera_position(camera_params.T_camera_world), unique_indexes=True)
  else:
    features = gaussians.feature[indexes]
    assert len(features.shape) == 2, f"Features must be (N, C) if use_sh=False, got {features.shape}"

  return render_projected(indexes, gaussians2d, features, depths, camera_params, config,
                          use_depth16=use_depth16, render_median_depth=render_median_depth, ndc_depths=ndc)


def render_projected(indexes: torch.Tensor, gaussians2d: torch.Tensor, features: torch.Tensor,
                     depths: torch.Tensor, camera_params: CameraParams, config: RasterConfig,
                     use_depth16: bool = False, render_median_depth: bool = False, ndc_depths=None) -> Rendering:
  if ndc_depths is None:
    ndc_depths = ndc_depth(depths.detach(), camera_params.near_plane, camera_params.far_plane)

  overlap_to_point, tile_overlap_ranges = map_to_tiles(gaussians2d, ndc_depths, image_size=camera_params.image_size,
                                                       config=config, use_depth16=use_depth16)
  ranges = tile_overlap_ranges.view(-1, 2)
  median_depth = None
  if render_median_depth:
    # the reference runs a second, non-blending raster pass over features=depths (:77-82); here the same value
    # comes out of the main pass (no gradient flows through it: the reference's non-blending backward is
    # disabled upstream)
    raster, median_depth = rasterize_with_tiles_and_median(
        gaussians2d, features, depths, overlap_to_point=overlap_to_point, tile_overlap_ranges=ranges,
        image_size=camera_params.image_size, config=config)
  else:
    raster = rasterize_with_tiles(gaussians2d, features, tile_overlap_ranges=ranges,
                                  overlap_to_point=overlap_to_point, image_size=camera_params.image_size, config=config)

  points = RenderedPoints(
      idx=indexes, depths=depths, gaussians2d=gaussians2d,
      _visibility=raster.visibility if config.compute_visibility else None,
      _prune_cost=raster.point_heuristic[:, 0] if config.compute_point_heuristic else None,
      _split_score=raster.point_heuristic[:, 1] if config.compute_point_heuristic else None,
      features=features, attributes=None, batch_size=(depths.shape[0],))

  return Rendering(image=raster.image, image_weight=raster.image_weight, depth_image=None,
                   median_depth_image=median_depth, points=points, camera=camera_params, config=config)


def viewspace_gradient(gaussians2d: torch.Tensor):
  assert gaussians2d.shape[1] == 7, f"Expected packed 2D gaussians (N,7), got {gaussians2d.shape}"
  assert gaussians2d.grad is not None, "Expected gradients on gaussians2d, run backward first with gaussians2d.retain_grad()"
  return torch.norm(gaussians2d.grad[:, :2], dim=1)
