"""N3 measurement: one visibility-aware optimiser step over the bench cloud (1 M Gaussians, SH degree 3, every point
visible) -- CUDA-event median, algorithmic bytes, fraction of the measured HBM peak -- and the CPU oracle on a bounded
sample beside it.  Usage: python profiles/bench_optim.py [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from taichi_splatting_b200 import optim

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = int(os.environ.get("GS_N", 1_000_000))
dev = torch.device("cuda:0")
SHAPES = {"position": ((3,), "vector"), "log_scaling": ((3,), "scalar"), "rotation": ((4,), "scalar"),
          "alpha_logit": ((1,), "scalar"), "feature": ((3, 16), "scalar")}


def make(device, count):
  gen = torch.Generator().manual_seed(0)
  params = {k: torch.randn((count, *s), generator=gen).to(device).requires_grad_(True) for k, (s, _) in SHAPES.items()}
  for p in params.values():
    p.grad = torch.randn(p.shape, generator=gen).to(device) if device == "cpu" else torch.randn_like(p)
  idx = torch.arange(count, device=device)
  vis = (torch.rand(count, generator=gen) * 3 + 0.01).to(device)
  return params, idx, vis


try:
  peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
  peak = 6650.0
for name in ("VisibilityAwareAdam", "VisibilityAwareLaProp", "SparseAdam"):
  params, idx, vis = make(dev, n)
  opt = getattr(optim, name)([dict(params=[p], name=k, type=SHAPES[k][1], lr=1e-3) for k, p in params.items()], lr=1e-3)
  call = (lambda: opt.step(idx)) if name.startswith("Sparse") else (lambda: opt.step(idx, vis))
  for _ in range(5):
    call()
  times = []
  for _ in range(steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    call()
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
  ms = sorted(times)[len(times) // 2]
  # per visible row: grad 4D read; first moment, parameter 8D read+write each; second moment 8D (scalar) | 8 (vector);
  # index 8, weight 4, total_weight 4 per group; visibility update 8 + 4 + 4 + 8 + 8 once
  d_scalar = sum(torch.Size(s).numel() for s, kind in SHAPES.values() if kind == "scalar")
  d_vector = sum(torch.Size(s).numel() for s, kind in SHAPES.values() if kind == "vector")
  per_row = 28 * d_scalar + 20 * d_vector + 8 + 16 * len(SHAPES) + (32 if name.startswith("Visibility") else 12)
  gbs = per_row * n / (ms * 1e-3) / 1e9
  print(f"{name}: {ms:.3f} ms per step over {n} visible Gaussians ({len(SHAPES)} parameter groups, {d_scalar + d_vector} floats / "
        f"Gaussian): {per_row} algorithmic B / Gaussian -> {gbs:.0f} GB/s = {gbs / peak:.2f} of the measured HBM peak ({peak:.0f} GB/s)")

# CPU oracle beside it (bounded sample)
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import optim_ops  # noqa: E402
m = 100_000
params, idx, vis = make("cpu", m)
state = {k: {} for k in params}
total, running = torch.zeros(m), torch.zeros(m)
groups = {k: dict(type=SHAPES[k][1], lr=1e-3, betas=(0.9, 0.999), eps=1e-16, bias_correction=True) for k in params}
t0 = time.perf_counter()
with torch.no_grad():
  w = optim_ops.update_visibility(running, vis, idx, 0.5)
  total[idx] += w
  for k, p in params.items():
    grad = p.grad.view(m, -1) / (vis.unsqueeze(1) + 0.01)
    optim_ops.group_step(groups[k], state[k], p.view(m, -1), grad, idx, w, total, optim_ops.ADAM)
dt = time.perf_counter() - t0
print(f"CPU oracle (torch, {torch.get_num_threads()} threads), VisibilityAwareAdam over {m} Gaussians: {dt * 1e3:.1f} ms "
      f"({m / dt / 1e6:.2f} M Gaussians/s; GPU {n / ms / 1e3:.0f} M Gaussians/s)")
