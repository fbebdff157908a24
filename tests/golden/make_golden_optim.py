"""Generates tests/golden/optim.npz: the REAL reference optimiser host code (optim/fractional.py,
optim/visibility_aware.py -- pure torch) run on CPU over the restated kernels of oracle/optim_ops.py.

  python tests/golden/make_golden_optim.py

Four parameter groups (vector, scalar + mask_lr, local_vector + basis, scalar + point_lr + clip), four sparse steps
with changing visible sets; parameters and every state tensor are recorded after each step, for each of
FractionalAdam, FractionalLaProp, SparseAdam, SparseLaProp, VisibilityAwareAdam, VisibilityAwareLaProp.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import optim_ops  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N, STEPS = 96, 4
GROUPS = [("position", "vector", (3,), 0.01), ("log_scaling", "scalar", (3,), 0.005),
          ("rotation", "local_vector", (3,), 0.002), ("feature", "scalar", (3, 4), 0.02)]
OPTIMISERS = ["FractionalAdam", "FractionalLaProp", "SparseAdam", "SparseLaProp", "VisibilityAwareAdam",
              "VisibilityAwareLaProp"]


def scenario(seed=0):
  gen = torch.Generator().manual_seed(seed)
  init = {name: torch.randn((N, *shape), generator=gen) for name, _, shape, _ in GROUPS}
  mask_lr = torch.tensor([1.0, 0.5, 2.0])
  point_lr = torch.rand(N, generator=gen) + 0.5
  steps = []
  for s in range(STEPS):
    m = int(torch.randint(N // 3, N, (1,), generator=gen))
    indexes = torch.randperm(N, generator=gen)[:m].sort().values
    steps.append(dict(indexes=indexes, weight=torch.rand(m, generator=gen) * 0.95 + 0.05,
                      visibility=torch.rand(m, generator=gen) * 5 + 0.01,
                      basis=torch.eye(3).unsqueeze(0) + 0.3 * torch.randn(m, 3, 3, generator=gen),
                      grads={name: torch.randn((N, *shape), generator=gen) for name, _, shape, _ in GROUPS}))
  return init, mask_lr, point_lr, steps


def param_groups(params, mask_lr, point_lr):
  groups = []
  for name, kind, _, lr in GROUPS:
    g = dict(params=[params[name]], name=name, type=kind, lr=lr)
    if name == "log_scaling":
      g["mask_lr"] = mask_lr
    if name == "feature":
      g["point_lr"], g["clip"] = point_lr, 0.5
    groups.append(g)
  return groups


def run(make_optimiser, kind, init, mask_lr, point_lr, steps, device="cpu"):
  """-> list over steps of {tensor name: array}; `make_optimiser(groups)`; kind: 'weight' | 'sparse' | 'visibility'."""
  params = {k: v.clone().to(device).requires_grad_(True) for k, v in init.items()}
  opt = make_optimiser(param_groups(params, mask_lr.to(device), point_lr.to(device)))
  out = []
  for st in steps:
    for k, p in params.items():
      p.grad = st["grads"][k].clone().to(device)
    idx, basis = st["indexes"].to(device), st["basis"].to(device)
    if kind == "sparse":
      opt.step(idx, basis=basis)
    elif kind == "visibility":
      opt.step(idx, st["visibility"].to(device), basis=basis)
    else:
      opt.step(idx, st["weight"].to(device), basis=basis)
    rec = {f"param_{k}": p.detach().cpu().numpy().copy() for k, p in params.items()}
    for k, p in params.items():
      for sk, sv in opt.state[p].items():
        if torch.is_tensor(sv):
          rec[f"state_{k}_{sk}"] = sv.detach().cpu().numpy().copy()
    out.append(rec)
  return out


def kind_of(name):
  return "sparse" if name.startswith("Sparse") else ("visibility" if name.startswith("Visibility") else "weight")


def main():
  fr, va = optim_ops.load_reference_optim()
  init, mask_lr, point_lr, steps = scenario()
  blob = {}
  for name in OPTIMISERS:
    cls = getattr(va if name.startswith("Visibility") else fr, name)
    recs = run(lambda groups: cls(groups, lr=0.01), kind_of(name), init, mask_lr, point_lr, steps)
    for s, rec in enumerate(recs):
      for k, v in rec.items():
        blob[f"{name}/{s}/{k}"] = v
  np.savez_compressed(os.path.join(HERE, "optim.npz"), **blob)
  print("wrote optim.npz:", len(blob), "arrays")


if __name__ == "__main__":
  main()
