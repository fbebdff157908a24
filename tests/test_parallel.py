"""Host logic of the multi-GPU path, run with world_size 2 on CPU (gloo).  The kernels themselves are covered
by the -m gpu tests; here: tile partitioning, gradient all-reduce (bucketed / per-tensor), and the
identity-forward / all-reduce-backward autograd node of the tile-sharded path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from taichi_splatting_b200 import parallel


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _worker(rank, world, port, results):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    assert parallel.world_info() == (rank, world)
    # view-parallel: sum of per-rank parameter gradients, bucketed and per-tensor, None grads count as zero
    for bucket in (True, False):
      torch.manual_seed(0)
      params = [torch.randn(50, 3, requires_grad=True), torch.randn(50, 1, requires_grad=True),
                torch.randn(50, 3, 16, requires_grad=True)]
      loss = sum(((r + 1) * (rank + 1)) * p.sum() for r, p in enumerate(params[:2]))
      loss.backward()          # params[2] gets no gradient on purpose
      parallel.allreduce_gradients(params, bucket=bucket)
      s = sum(k + 1 for k in range(world))
      assert torch.allclose(params[0].grad, torch.full((50, 3), 1.0 * s))
      assert torch.allclose(params[1].grad, torch.full((50, 1), 2.0 * s))
      assert torch.equal(params[2].grad, torch.zeros(50, 3, 16))
    # tile-sharded: each rank sees only its tiles' loss; gradients of the shared inputs are summed
    x = torch.arange(6, dtype=torch.float64).requires_grad_(True)
    y = torch.ones(4, dtype=torch.float64, requires_grad=True)
    xr, yr = parallel.reduce_across_ranks(x, y)
    ((rank + 1) * (xr * xr).sum() + (rank + 2) * yr.sum()).backward()
    assert torch.allclose(x.grad, 2 * x.detach() * sum(k + 1 for k in range(world)))
    assert torch.allclose(y.grad, torch.full((4,), float(sum(k + 2 for k in range(world))), dtype=torch.float64))
    assert parallel.views_for_rank(8, rank, world) == list(range(rank, 8, world))
    # tile shard: equal tile counts first, then equal overlap counts from the ranks' own (disjoint) tile ranges
    shard = parallel.TileShard()
    T = 40
    lo, hi = shard.tile_range(T)
    assert (lo, hi) == (T * rank // world, T * (rank + 1) // world)
    counts = torch.arange(T, dtype=torch.int32) * 3 + 1             # heavier tiles at the end of the grid
    ends = torch.cumsum(counts, 0).to(torch.int32)
    full = torch.stack([ends - counts, ends], 1)
    mine = torch.zeros_like(full)
    mine[lo:hi] = full[lo:hi] - full[lo, 0]                         # a rank's ranges index its OWN overlap list
    b = shard.rebalance(mine.view(4, 10, 2))
    assert b.tolist() == parallel.partition_tiles(full, world).tolist() and shard.tile_range(T) == (int(b[rank]), int(b[rank + 1]))
    per_rank = [int(counts[int(b[r]):int(b[r + 1])].sum()) for r in range(world)]
    assert max(per_rank) - min(per_rank) <= int(counts.max())
    flat = torch.full((6,), float(rank + 1))
    shard.reduce(flat)
    assert torch.equal(flat, torch.full((6,), float(sum(k + 1 for k in range(world)))))
    results[rank] = "ok"
  finally:
    dist.destroy_process_group()


def test_world_size_2_gloo():
  world = 2
  port = _free_port()
  ctx = mp.get_context("spawn")
  results = ctx.Manager().dict()
  procs = [ctx.Process(target=_worker, args=(r, world, port, results)) for r in range(world)]
  for p in procs:
    p.start()
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  assert dict(results) == {0: "ok", 1: "ok"}


def test_partition_tiles_balances_overlaps():
  torch.manual_seed(0)
  counts = torch.randint(0, 500, (34, 60))
  counts[5:9] = 0                                     # empty rows of tiles
  ends = torch.cumsum(counts.reshape(-1), 0)
  starts = ends - counts.reshape(-1)
  ranges = torch.stack([starts, ends], 1).to(torch.int32)
  ranges[counts.reshape(-1) == 0] = 0                 # the mapper leaves untouched tiles at (0, 0)
  total = int(counts.sum())
  for world in (1, 2, 4, 8):
    b = parallel.partition_tiles(ranges.view(34, 60, 2), world)
    assert b.shape == (world + 1,) and int(b[0]) == 0 and int(b[-1]) == 34 * 60
    assert bool((b[1:] >= b[:-1]).all())
    per_rank = [int(counts.reshape(-1)[int(b[r]):int(b[r + 1])].sum()) for r in range(world)]
    assert sum(per_rank) == total
    assert max(per_rank) - total / world <= 500, per_rank     # off by at most one tile's worth
    covered = torch.zeros(34 * 60, dtype=torch.int32)
    for r in range(world):
      m = parallel.mask_tile_ranges(ranges, int(b[r]), int(b[r + 1]))
      covered += ((m[:, 1] - m[:, 0]) > 0).int()
      assert torch.equal(m[int(b[r]):int(b[r + 1])], ranges[int(b[r]):int(b[r + 1])])
    assert torch.equal(covered, (counts.reshape(-1) > 0).int())   # every non-empty tile rendered exactly once
  b = parallel.partition_tiles(torch.zeros((16, 2), dtype=torch.int32), 4)
  assert b.tolist() == [0, 4, 8, 12, 16]
