"""taichi_splatting_b200 -- B200-native (sm_100a) drop-in for taichi-splatting's render hot path.

Same operator surface as `taichi_splatting` (reference __init__.py:1-33); every operator is a
torch.autograd.Function over the C-ABI CUDA library libgsplat_b200.so.  No Taichi, no Triton, no CPU path.
"""
from . import perspective
from .data_types import Gaussians2D, Gaussians3D, RasterConfig
from .mapper.tile_mapper import map_to_tiles, pad_to_tile
from .rasterizer import RasterOut, rasterize, rasterize_with_tiles
from .renderer import render_gaussians, render_projected
from .rendering import Rendering, RenderedPoints
from .spherical_harmonics import evaluate_sh, evaluate_sh_at
from .taichi_queue import TaichiQueue

__all__ = [
    'render_gaussians', 'render_projected', 'Rendering', 'RenderedPoints',
    'map_to_tiles', 'pad_to_tile',
    'Gaussians2D', 'Gaussians3D', 'RasterConfig',
    'evaluate_sh', 'evaluate_sh_at',
    'rasterize', 'rasterize_with_tiles', 'RasterOut',
    'perspective', 'TaichiQueue',
]
