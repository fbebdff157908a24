"""How far the image-fitting loop with densification ends above the fixed cloud, over repeated runs (the float atomics
of the raster backward make every run slightly different): picks the configuration of the GPU test."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import contextlib, io
from taichi_splatting_b200.examples import fit_image_gaussians as ex

for detail, n, target in ((1.0, 600, 2400), (3.0, 150, 1500), (4.0, 200, 2400)):
  for rep in range(3):
    with contextlib.redirect_stdout(io.StringIO()):
      fixed = ex.main(["--n", str(n), "--iters", "300", "--size", "192,160", "--detail", str(detail)])
      grown = ex.main(["--n", str(n), "--iters", "300", "--size", "192,160", "--detail", str(detail), "--target", str(target), "--epoch", "50"])
    print(f"detail {detail} n {n} -> {target}: fixed {fixed:.2f} dB, grown {grown:.2f} dB, margin {grown - fixed:+.2f}")
