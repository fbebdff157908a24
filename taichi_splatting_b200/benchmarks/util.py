"""Timing protocol of the benchmark CLIs (reference: taichi_splatting/benchmarks/util.py:23-37 -- warm-up iterations,
CUDA events around the loop, one synchronise at the end).  `--profile` prints the torch.profiler kernel table instead of
Taichi's kernel profiler."""
import torch


def benchmarked(name: str, f, iters: int = 100, warmup: int = 10, profile: bool = False, quiet: bool = False) -> float:
  """Runs f() `warmup` + `iters` times; prints and returns iterations per second (device time, CUDA events)."""
  for _ in range(warmup):
    f()
  if profile:
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
      for _ in range(min(iters, 20)):
        f()
      torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12))
  start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  start.record()
  for _ in range(iters):
    f()
  end.record()
  torch.cuda.synchronize()
  elapsed = start.elapsed_time(end) / 1000.
  rate = iters / elapsed
  if not quiet:
    print(f"{name}  {iters} iterations in {elapsed:.3f}s at {rate:.1f} iters/sec ({1e3 * elapsed / iters:.4f} ms)")
  return rate


def size_arg(text: str):
  return tuple(map(int, text.split(",")))
