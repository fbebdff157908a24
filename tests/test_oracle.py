"""CPU tests that PIN the oracle (no GPU).

 * torch_ops.project / evaluate_sh_at  vs golden vectors produced by the REAL reference torch_lib
   (tests/golden/make_golden.py), and vs the live reference import when /root/reference exists.
 * C rasteriser: backward == finite differences of forward (fp64) -- the property the reference's
   own gradcheck test pins (tests/test_rasterizer.py:84-90); visibility == d(sum image)/d feature
   (tests/test_visibility.py:34-64); stale-group emulation == intended semantics when every tile
   holds <= 64 overlaps.
 * C tile mapper: brute-force geometric check of the OBB test, sortedness, range consistency.
"""
import numpy as np
import pytest
import torch

from oracle import cbind, random_data, ref_loader, torch_ops
from oracle.cbind import OracleConfig

PROJ_NAMES = ["position", "log_scaling", "rotation", "alpha_logit", "T_camera_world", "projection"]


def _proj_cases(golden_dir):
  z = np.load(f"{golden_dir}/projection.npz")
  prefixes = sorted({"_".join(k.split("_")[:2]) for k in z.files})
  for p in prefixes:
    yield p, {k[len(p) + 1:]: z[k] for k in z.files if k.startswith(p + "_")}


def test_projection_matches_reference_golden(golden_dir):
  n_cases = 0
  for name, c in _proj_cases(golden_dir):
    ins = [torch.from_numpy(c[f"in_{k}"]).clone().requires_grad_(True) for k in PROJ_NAMES]
    pts, depth, idx = torch_ops.project(*ins, tuple(int(v) for v in c["image_size"]),
                                        tuple(float(v) for v in c["depth_range"]), blur_cov=float(c["blur_cov"]))
    assert np.array_equal(idx.numpy(), c["indexes"]), name
    f64 = pts.dtype == torch.float64
    tol = dict(rtol=1e-9, atol=1e-11) if f64 else dict(rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(pts.detach().numpy(), c["points"], err_msg=name, **tol)
    np.testing.assert_allclose(depth.detach().numpy(), c["depth"], err_msg=name, **tol)
    (pts.mean() + depth.mean()).backward()
    for k, t in zip(PROJ_NAMES, ins):
      ref = c[f"grad_{k}"]
      gt = dict(rtol=1e-7, atol=1e-9 * max(1.0, np.abs(ref).max())) if f64 else \
           dict(rtol=2e-3, atol=2e-4 * max(1.0, np.abs(ref).max()))
      np.testing.assert_allclose(t.grad.numpy(), ref, err_msg=f"{name} grad {k}", **gt)
    n_cases += 1
  assert n_cases == 12


def test_sh_matches_reference_golden(golden_dir):
  z = np.load(f"{golden_dir}/spherical_harmonics.npz")
  for seed in range(12):
    g = lambda k: z[f"sh_{seed}_{k}"]
    params = torch.from_numpy(g("in_params")).clone().requires_grad_(True)
    points = torch.from_numpy(g("in_points")).clone().requires_grad_(True)
    cam = torch.from_numpy(g("in_camera_pos")).clone().requires_grad_(True)
    out = torch_ops.evaluate_sh_at(params, points, torch.from_numpy(g("indexes")), cam)
    tol = dict(rtol=1e-9, atol=1e-12) if out.dtype == torch.float64 else dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out.detach().numpy(), g("out"), **tol)
    out.mean().backward()
    np.testing.assert_allclose(params.grad.numpy(), g("grad_params"), **tol)
    np.testing.assert_allclose(points.grad.numpy(), g("grad_points"), **tol)
    np.testing.assert_allclose(cam.grad.numpy(), g("grad_camera_pos"), **tol)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_projection_matches_live_reference():
  proj, sh = ref_loader.load()
  for seed in range(20, 26):
    torch.manual_seed(seed)
    cam = random_data.random_camera()
    g = random_data.random_3d_gaussians(500, cam, margin=0.5, scale_factor=0.5)
    args = [x.double() for x in (g.position, g.log_scaling, g.rotation, g.alpha_logit,
                                 cam.T_camera_world, cam.projection)]
    a = proj.apply(*args, cam.image_size, cam.depth_range, blur_cov=0.3)
    b = torch_ops.project(*args, cam.image_size, cam.depth_range, blur_cov=0.3)
    assert torch.equal(a[2], b[2])
    assert torch.allclose(a[0], b[0], rtol=1e-10, atol=1e-12) and torch.allclose(a[1], b[1])
    nd_a = (1 - (1. / a[1] - 1. / cam.far_plane) / (1. / cam.near_plane - 1. / cam.far_plane))
    assert torch.allclose(nd_a, torch_ops.ndc_depth(b[1], cam.near_plane, cam.far_plane))


def _single_tile_case(seed, ts=8, dtype=np.float64):
  # tests/test_rasterizer.py:30-59 make_inputs: one ts x ts tile, identity overlap list
  torch.random.manual_seed(seed)
  n = torch.randint(1, 50, (1,)).item()
  channels = torch.randint(1, 4, (1,)).item()
  g = random_data.random_2d_gaussians(n, (ts, ts), num_channels=channels, scale_factor=1.0, alpha_range=(0.2, 0.8))
  pts = random_data.packed_2d(g).numpy().astype(dtype)
  feat = g.feature.numpy().astype(dtype)
  o2p = np.arange(n, dtype=np.int32)
  ranges = np.array([[0, n]], np.int32)
  return pts, feat, o2p, ranges


@pytest.mark.parametrize("antialias", [False, True])
def test_raster_backward_is_derivative_of_forward(antialias):
  cfg = OracleConfig(tile_size=8, pixel_stride=(1, 1), antialias=antialias, saturate_threshold=2.0)
  rng = np.random.default_rng(0)
  for seed in range(6):
    pts, feat, o2p, ranges = _single_tile_case(seed)
    R = rng.uniform(size=(8, 8, feat.shape[1]))
    image, _, _ = cbind.raster_forward(pts, feat, ranges, o2p, (8, 8), cfg, dtype=np.float64)
    gp, gf, _ = cbind.raster_backward(pts, feat, ranges, o2p, image, R, (8, 8), cfg, dtype=np.float64)

    def loss(p, f):
      return float((cbind.raster_forward(p, f, ranges, o2p, (8, 8), cfg, dtype=np.float64)[0] * R).sum())

    eps = 1e-6
    for arr, grad, is_pts in ((pts, gp, True), (feat, gf, False)):
      num = np.zeros_like(arr)
      for i in np.ndindex(arr.shape):
        a, b = arr.copy(), arr.copy()
        a[i] += eps
        b[i] -= eps
        num[i] = (loss(a, feat) - loss(b, feat)) / (2 * eps) if is_pts else (loss(pts, a) - loss(pts, b)) / (2 * eps)
      np.testing.assert_allclose(grad, num, rtol=2e-5, atol=2e-7, err_msg=f"seed {seed}")


def test_visibility_equals_feature_gradient():
  # tests/test_visibility.py:34-64 (320x200, loss = image.sum()), with the saturation skip disabled
  cfg = OracleConfig(compute_visibility=True, compute_point_heuristic=True, saturate_threshold=2.0)
  for seed in range(3):
    torch.manual_seed(seed)
    n = int(np.random.default_rng(seed).integers(1, 4000))
    g = random_data.random_2d_gaussians(n, (320, 200), scale_factor=0.2, alpha_range=(0.2, 1.0))
    pts = random_data.packed_2d(g).numpy().astype(np.float64)
    feat = g.feature.numpy().astype(np.float64)
    depth = g.depths.clamp(0, 1).numpy().astype(np.float32)
    o2p, ranges = cbind.map_to_tiles(pts.astype(np.float32), depth, (320, 200), cfg)
    image, alpha, vis = cbind.raster_forward(pts, feat, ranges, o2p, (320, 200), cfg, dtype=np.float64)
    gp, gf, heur = cbind.raster_backward(pts, feat, ranges, o2p, image, np.ones_like(image), (320, 200), cfg, dtype=np.float64)
    np.testing.assert_allclose(gf[:, 0], vis, rtol=1e-9, atol=1e-12)
    assert heur.shape == (n, 2) and (heur >= 0).all()


def test_stale_group_emulation_is_identity_for_small_tiles():
  cfg = OracleConfig(tile_size=16)
  torch.manual_seed(3)
  g = random_data.random_2d_gaussians(300, (64, 64), scale_factor=0.5)
  pts = random_data.packed_2d(g).numpy()
  o2p, ranges = cbind.map_to_tiles(pts, g.depths.numpy(), (64, 64), cfg)
  assert (ranges[..., 1] - ranges[..., 0]).max() <= 64
  a = cbind.raster_forward(pts, g.feature.numpy(), ranges, o2p, (64, 64), cfg)[0]
  b = cbind.raster_forward(pts, g.feature.numpy(), ranges, o2p, (64, 64), cfg, emulate_stale_group=True)[0]
  assert np.array_equal(a, b)
  # ... and differs (D1) once a tile holds more than one group
  g = random_data.random_2d_gaussians(6000, (64, 64), scale_factor=2.0)
  pts = random_data.packed_2d(g).numpy()
  o2p, ranges = cbind.map_to_tiles(pts, g.depths.numpy(), (64, 64), cfg)
  assert (ranges[..., 1] - ranges[..., 0]).max() > 256
  a = cbind.raster_forward(pts, g.feature.numpy(), ranges, o2p, (64, 64), cfg)[0]
  b = cbind.raster_forward(pts, g.feature.numpy(), ranges, o2p, (64, 64), cfg, emulate_stale_group=True)[0]
  assert not np.array_equal(a, b)


def _brute_force_overlap(g, tx, ty, ts, thr, samples=24):
  """Does the alpha>thr ellipse of g touch tile (tx,ty)?  Dense sampling of the tile (conservative: only
  used to assert that every tile containing an above-threshold sample IS reported by the OBB test)."""
  xs = (np.arange(samples) + 0.5) / samples * ts
  X, Y = np.meshgrid(tx * ts + xs, ty * ts + xs)
  dx, dy = X - g[0], Y - g[1]
  u = (dx * g[2] + dy * g[3]) / g[4]
  v = (-dx * g[3] + dy * g[2]) / g[5]
  return bool((g[6] * np.exp(-0.5 * (u * u + v * v)) > thr).any())


def test_tile_mapper_geometry_and_order():
  cfg = OracleConfig(tile_size=16)
  size = (200, 120)  # not a multiple of the tile size on purpose
  torch.manual_seed(7)
  g = random_data.random_2d_gaussians(800, size, scale_factor=1.5, alpha_range=(0.05, 0.95))
  pts = random_data.packed_2d(g).numpy()
  depth = g.depths.numpy()
  o2p, ranges, keys, counts = cbind.map_to_tiles(pts, depth, size, cfg, return_keys=True)
  TH, TW = ranges.shape[:2]
  assert (TH, TW) == (8, 13) and counts.sum() == len(o2p) == len(keys)
  # keys sorted on the low 48 bits; ties keep ascending gaussian index (stable sort)
  assert (np.diff(keys.astype(np.int64)) >= 0).all()
  same = np.diff(keys.astype(np.int64)) == 0
  assert (np.diff(o2p)[same] > 0).all()
  # ranges partition the overlap list by tile id; depth ascending inside a tile
  tile_of = (keys >> np.uint64(32)).astype(np.int64)
  flat = ranges.reshape(-1, 2)
  covered = 0
  for t in range(TH * TW):
    s, e = flat[t]
    if e > s:
      assert (tile_of[s:e] == t).all()
      d = depth[o2p[s:e], 0]
      assert (np.diff(d) >= 0).all()
      covered += e - s
    else:
      assert s == 0 and e == 0
  assert covered == len(o2p)
  # every tile with an above-threshold sample must be reported for that gaussian (no false negatives)
  reported = {(int(p), int(t)) for p, t in zip(o2p, tile_of)}
  for i in range(0, 800, 7):
    for ty in range(TH):
      for tx in range(TW):
        if _brute_force_overlap(pts[i].astype(np.float64), tx, ty, 16, cfg.alpha_threshold):
          assert (i, tx + ty * TW) in reported, (i, tx, ty)


def test_tile_mapper_edge_cases():
  cfg = OracleConfig()
  o2p, ranges = cbind.map_to_tiles(np.zeros((0, 7), np.float32), np.zeros((0, 1), np.float32), (64, 48), cfg)
  assert o2p.shape == (0,) and ranges.shape == (3, 4, 2) and not ranges.any()
  # alpha below the threshold -> no tiles; fully off-image gaussian -> no tiles
  pts = np.array([[10, 10, 1, 0, 3, 3, 0.001], [-500, -500, 1, 0, 3, 3, 0.9], [30, 20, 0.6, 0.8, 5, 2, 0.9]], np.float32)
  counts = cbind.tile_counts(pts, (64, 48), cfg)
  assert counts[0] == 0 and counts[2] > 0
  o2p, ranges = cbind.map_to_tiles(pts, np.array([[0.5], [0.2], [0.7]], np.float32), (64, 48), cfg)
  assert set(o2p.tolist()) <= {1, 2}
  # depth16 keys: same tile partition, order by quantised depth then index
  torch.manual_seed(1)
  g = random_data.random_2d_gaussians(500, (64, 48), scale_factor=1.0)
  p = random_data.packed_2d(g).numpy()
  o32, r32 = cbind.map_to_tiles(p, g.depths.numpy(), (64, 48), cfg)
  o16, r16 = cbind.map_to_tiles(p, g.depths.numpy(), (64, 48), cfg, use_depth16=True)
  assert np.array_equal(r32, r16) and sorted(o32.tolist()) == sorted(o16.tolist())


def test_morton_oracle_bit_interleaving():
  """N4: the spreading trick of the Morton oracle against a bit-by-bit interleave; codes sort cells in Z-order."""
  from oracle import morton
  rng = np.random.default_rng(0)
  cells = rng.integers(0, 2**21, size=(200, 3), dtype=np.uint64)
  fast = morton.spread_bits64(cells[:, 0]) | (morton.spread_bits64(cells[:, 1]) << np.uint64(1)) | (morton.spread_bits64(cells[:, 2]) << np.uint64(2))
  slow = np.array([morton.interleave_slow(int(a), int(b), int(c)) for a, b, c in cells], dtype=np.uint64)
  assert np.array_equal(fast, slow)
  pts = rng.uniform(-5, 5, size=(1000, 3)).astype(np.float32)
  order = morton.argsort(pts, 0.01)
  codes = morton.morton_codes64(pts, 0.01)
  assert np.all(np.diff(codes[order].astype(np.float64)) >= 0) and sorted(order.tolist()) == list(range(1000))
