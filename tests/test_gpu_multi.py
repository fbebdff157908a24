"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): torchrun-style workers over NCCL.

 * tile-sharded single view: union of the per-rank image strips == single-GPU image, and every rank ends with the
   single-GPU parameter gradients (the one all-reduce of packed-2D / feature gradients).
 * view-parallel: all-reduced gradients == sum of the per-view single-GPU gradients.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _scene(ts, dev, yaw=0.0):
  from taichi_splatting_b200.benchmarks import scenes
  size = (320, 208)
  cam = scenes.benchmark_camera(size, yaw_deg=yaw).to(device=dev)
  cloud = scenes.random_3d_gaussians(30000, scenes.benchmark_camera(size), scale_factor=1.5, sh_degree=1, seed=3)
  return cloud.to(dev).requires_grad_(True), cam, size


def _rel(a, b):
  return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _worker(rank, world, port, results):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.cuda.set_device(rank)
  dev = torch.device("cuda", rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
  try:
    import taichi_splatting_b200 as ts
    from taichi_splatting_b200 import parallel
    cfg = ts.RasterConfig()
    names = ("position", "log_scaling", "rotation", "alpha_logit", "feature")
    R = torch.rand((208, 320, 3), generator=torch.Generator().manual_seed(0)).to(dev)

    # ---- single-GPU reference on every rank ----
    cloud, cam, size = _scene(ts, dev)
    ref = ts.render_gaussians(cloud, cam, cfg, use_sh=True)
    (ref.image * R).sum().backward()
    ref_grads = {k: getattr(cloud, k).grad.clone() for k in names}

    # ---- tile-sharded: whole-frame drivers (each rank bins / sorts / packs / rasterises its own tile range), first with
    # equal tile counts, then rebalanced to equal overlap counts; then the operator-chain form (tile size 8) ----
    shard = parallel.TileShard()
    ks = []
    for frame in range(2):
      cloud2, cam2, _ = _scene(ts, dev)
      out, (lo, hi) = parallel.render_tile_sharded(cloud2, cam2, cfg, use_sh=True, shard=shard)
      (out.image * R).sum().backward()
      full = out.image.detach().clone()
      dist.all_reduce(full)                       # strips are disjoint, so the sum is the union
      assert _rel(full, ref.image.detach()) < 1e-6, _rel(full, ref.image.detach())
      for k in names:
        assert _rel(getattr(cloud2, k).grad, ref_grads[k]) < 2e-5, (frame, k, _rel(getattr(cloud2, k).grad, ref_grads[k]))
      assert 0 <= lo <= hi
      k_all = torch.tensor([shard.last_k], device=dev)
      gathered = [torch.zeros_like(k_all) for _ in range(world)]
      dist.all_gather(gathered, k_all)
      ks.append([int(x) for x in gathered])
      shard.rebalance()
    assert sum(ks[0]) == sum(ks[1])                                            # same overlaps, cut differently
    assert max(ks[1]) - min(ks[1]) <= max(ks[0]) - min(ks[0]) + 600, ks         # rebalanced: no worse than equal tiles
    cfg8 = ts.RasterConfig(tile_size=8)
    ref8_cloud, ref8_cam, _ = _scene(ts, dev)
    ref8 = ts.render_gaussians(ref8_cloud, ref8_cam, cfg8, use_sh=True)
    (ref8.image * R).sum().backward()
    cloud3, cam3, _ = _scene(ts, dev)
    out8, _ = parallel.render_tile_sharded(cloud3, cam3, cfg8, use_sh=True)
    (out8.image * R).sum().backward()
    full8 = out8.image.detach().clone()
    dist.all_reduce(full8)
    assert _rel(full8, ref8.image.detach()) < 1e-6
    for k in names:
      assert _rel(getattr(cloud3, k).grad, getattr(ref8_cloud, k).grad) < 2e-5, ("operators", k)

    # ---- view-parallel ----
    per_view = []
    for r in range(world):
      c, cm, _ = _scene(ts, dev, yaw=2.0 * r)
      o = ts.render_gaussians(c, cm, cfg, use_sh=True)
      (o.image * R).sum().backward()
      per_view.append({k: getattr(c, k).grad for k in names})
    c, cm, _ = _scene(ts, dev, yaw=2.0 * rank)
    o = ts.render_gaussians(c, cm, cfg, use_sh=True)
    (o.image * R).sum().backward()
    parallel.allreduce_gradients([getattr(c, k) for k in names])
    for k in names:
      want = sum(pv[k] for pv in per_view)
      assert _rel(getattr(c, k).grad, want) < 2e-5, (k, _rel(getattr(c, k).grad, want))
    # same sums through the structured exchange (SH gradient via its rank-1 factors, geometry via all-reduce)
    c, cm, _ = _scene(ts, dev, yaw=2.0 * rank)
    o = parallel.render_view_parallel(c, cm, cfg, use_sh=True)
    (o.image * R).sum().backward()
    parallel.finish_view_parallel_backward(c, use_sh=True)
    for k in names:
      want = sum(pv[k] for pv in per_view)
      assert _rel(getattr(c, k).grad, want) < 2e-5, ("structured", k, _rel(getattr(c, k).grad, want))
    # ... with everything reduced inside the backward, over NCCL and over peer memory (fused pack + all-gather kernel,
    # in-place peer all-reduce); several frames, so that both gathered slots and the persistent buffers are reused
    for mode in ("0", "1"):
      os.environ["GS_PEER_EXCHANGE"] = mode
      for frame in range(4):
        c, cm, _ = _scene(ts, dev, yaw=2.0 * rank)
        o = parallel.render_view_parallel(c, cm, cfg, use_sh=True, reduce_in_backward=True)
        (o.image * R).sum().backward()
        for k in names:
          want = sum(pv[k] for pv in per_view)
          assert _rel(getattr(c, k).grad, want) < 2e-5, ("in-backward", mode, frame, k, _rel(getattr(c, k).grad, want))
    os.environ.pop("GS_PEER_EXCHANGE", None)
    results[rank] = "ok"
  finally:
    dist.destroy_process_group()


def test_two_gpu_sharding_matches_single_gpu():
  if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs >= 2 CUDA devices")
  world, port = 2, _free_port()
  ctx = mp.get_context("spawn")
  results = ctx.Manager().dict()
  procs = [ctx.Process(target=_worker, args=(r, world, port, results)) for r in range(world)]
  for p in procs:
    p.start()
  for p in procs:
    p.join(timeout=600)
    assert p.exitcode == 0
  assert dict(results) == {0: "ok", 1: "ok"}


def test_second_device_in_one_process():
  """A single process driving two GPUs: tensors on cuda:1 while cuda:0 is the current device.  Kernel attributes
  (the backward's > 48 KB dynamic shared memory), library scratch and the auxiliary streams are per device, so the
  call must run on the tensors' device (_lib.call switches to it) -- this failed with 'invalid argument' before."""
  if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs >= 2 CUDA devices")
  import taichi_splatting_b200 as ts
  torch.cuda.set_device(0)
  cfg = ts.RasterConfig(compute_visibility=True, compute_point_heuristic=True)
  R = torch.rand((208, 320, 3), generator=torch.Generator().manual_seed(0))
  grads, images = [], []
  for dev in ("cuda:0", "cuda:1", "cuda:0"):
    cloud, cam, _ = _scene(ts, torch.device(dev))
    out = ts.render_gaussians(cloud, cam, cfg, use_sh=True, render_median_depth=True)
    (out.image * R.to(dev)).sum().backward()
    assert torch.cuda.current_device() == 0
    grads.append(cloud.position.grad.cpu())
    images.append(out.image.detach().cpu())
  assert torch.equal(images[0], images[1]) and torch.equal(images[0], images[2])
  assert _rel(grads[1], grads[0]) < 1e-5 and _rel(grads[2], grads[0]) < 1e-5
