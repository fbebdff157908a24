"""CPU oracle of the sparse / visibility-aware optimisers (N3).  TEST INFRASTRUCTURE ONLY.

* `scalar_kernel` / `vector_kernel`: torch restatements of the reference's four Taichi kernels
  (optim/fractional_adam.py:7-44,46-86, optim/fractional_laprop.py:6-39,41-76), fp32, same operation order, with the
  reference's call signature -- so the REAL reference host code (optim/fractional.py, optim/visibility_aware.py, pure
  torch) can run on top of them: `load_reference_optim()` does that where /root/reference exists and
  tests/golden/make_golden_optim.py records the result.  The kernels themselves are Taichi and cannot run here
  (no taichi wheel): that part is restated, not pinned.
* `fractional_step` / `visibility_step`: restatement of the host logic (fractional.py:109-199,
  visibility_aware.py:37-105), checked against those golden vectors by the CPU tests.
"""
import importlib.util
import os
import sys
import types

import torch

ADAM, LAPROP = 0, 1


def _lerp(t, a, b):   # taichi_lib/generic.py:489-490
  return a * t + b * (1.0 - t)


def scalar_kernel(algorithm, betas=(0.9, 0.999), eps=1e-16, bias_correction=True):
  beta1, beta2 = (torch.tensor(b, dtype=torch.float32) for b in betas)

  def kernel(lr_step, indexes, weight, m_arr, v_arr, total_weight, grad, lr):
    w = weight.to(torch.float32).unsqueeze(1)
    tw = total_weight[indexes].unsqueeze(1)
    g = grad[indexes]
    if algorithm == ADAM:      # fractional_adam.py:26-41
      bias = torch.sqrt(1 - beta2 ** tw) / (1 - beta1 ** tw) if bias_correction else 1.0
      m = _lerp(beta1 ** w, m_arr[indexes], g)
      v = _lerp(beta2 ** w, v_arr[indexes], g * g)
      lr_step[:] = m / torch.clamp_min(torch.sqrt(v), eps) * bias * lr
    else:                      # fractional_laprop.py:22-37
      bias1 = 1.0 - beta1 ** tw if bias_correction else 1.0
      bias2 = 1.0 - beta2 ** tw if bias_correction else 1.0
      v = _lerp(beta2 ** w, v_arr[indexes], g * g)
      m = _lerp(beta1 ** w, m_arr[indexes], g / torch.clamp_min(torch.sqrt(v / bias2), eps))
      lr_step[:] = m * lr / bias1
    m_arr[indexes] = m
    v_arr[indexes] = v
  return kernel


def vector_kernel(algorithm, betas=(0.9, 0.999), eps=1e-16, dims=3, bias_correction=True):
  beta1, beta2 = (torch.tensor(b, dtype=torch.float32) for b in betas)

  def kernel(lr_step, indexes, weight, m_arr, v_arr, total_weight, grad, lr):
    w = weight.to(torch.float32)
    tw = total_weight[indexes]
    g = grad[indexes]
    norm = (g * g).sum(dim=1)
    v = _lerp(beta2 ** w, v_arr[indexes], norm)
    if algorithm == ADAM:      # fractional_adam.py:69-84
      bias = torch.sqrt(1 - beta2 ** tw) / (1 - beta1 ** tw) if bias_correction else torch.ones_like(tw)
      m = _lerp((beta1 ** w).unsqueeze(1), m_arr[indexes], g)
      lr_step[:] = (m / torch.clamp_min(torch.sqrt(v), eps).unsqueeze(1)) * (bias * lr).unsqueeze(1)
    else:                      # fractional_laprop.py:60-74
      bias1 = 1.0 - beta1 ** tw if bias_correction else torch.ones_like(tw)
      bias2 = 1.0 - beta2 ** tw if bias_correction else torch.ones_like(tw)
      m = _lerp((beta1 ** w).unsqueeze(1), m_arr[indexes], g / torch.clamp_min(torch.sqrt(v / bias2), eps).unsqueeze(1))
      lr_step[:] = m * (lr / bias1).unsqueeze(1)
    m_arr[indexes] = m
    v_arr[indexes] = v
  return kernel


def kernel_module(algorithm):
  """A stand-in for the reference's fractional_adam / fractional_laprop modules (same factory signatures)."""
  m = types.ModuleType("fractional_laprop" if algorithm == LAPROP else "fractional_adam")
  m.scalar_kernel = lambda betas=(0.9, 0.999), eps=1e-16, bias_correction=True: scalar_kernel(algorithm, betas, eps, bias_correction)
  m.vector_kernel = lambda betas=(0.9, 0.999), eps=1e-16, dims=3, bias_correction=True: vector_kernel(algorithm, betas, eps, dims, bias_correction)
  return m


# ---------------------------------------------------------------------------------- restated host logic
def saturate(x):   # fractional.py:149-150
  return 1 - 1 / torch.exp(2 * x)


def _state(state, param, vector):   # util.py:5-19 (note the reference's naming: first moment under 'v')
  if 'v' not in state:
    state['v'] = torch.zeros_like(param)
    state['m'] = torch.zeros((param.shape[0],), dtype=param.dtype) if vector else torch.zeros_like(param)
  return state['v'], state['m']


def group_step(g, state, param, grad, indexes, weight, total_weight, algorithm, basis=None):
  """One group (dict with type, lr, betas, eps, bias_correction, clip, mask_lr, point_lr): fractional.py:109-147,197-199."""
  vector = g["type"] in ("vector", "local_vector")
  first, second = _state(state, param, vector)
  make = vector_kernel if vector else scalar_kernel
  kw = dict(dims=param.shape[1]) if vector else {}
  kernel = make(algorithm, betas=g["betas"], eps=g["eps"], bias_correction=g["bias_correction"], **kw)
  if g["type"] == "local_vector":
    grad = grad.clone()
    grad[indexes] = torch.einsum('bij,bj->bi', torch.linalg.inv(basis), grad[indexes])
  lr_step = param.new_zeros(indexes.shape[0], param.shape[1])
  kernel(lr_step, indexes, weight, first, second, total_weight, grad, g["lr"])
  if g.get("clip") is not None:
    lr_step.clamp_(-g["lr"] * g["clip"], g["lr"] * g["clip"])
  if g["type"] == "local_vector":
    lr_step = torch.einsum('bij,bj->bi', basis, lr_step)
  if g.get("mask_lr") is not None:
    lr_step *= g["mask_lr"].view(-1).unsqueeze(0)
  if g.get("point_lr") is not None:
    lr_step *= g["point_lr"][indexes].unsqueeze(1)
  lr_step[~lr_step.isfinite()] = 0.0
  param[indexes] -= lr_step * saturate(weight).unsqueeze(1)


def update_visibility(running_vis, visibility, indexes, beta, eps=1e-12):   # visibility_aware.py:26-48
  a, b = visibility ** 4, running_vis[indexes] ** 4
  updated = (a + (b - a) * beta) ** 0.25
  running_vis[indexes] = updated
  return visibility / torch.clamp_min(updated, eps)


# ---------------------------------------------------------------------------------- the reference's host code
REF = os.environ.get("GS_REFERENCE_ROOT", "/root/reference") + "/taichi_splatting"


def reference_available():
  return os.path.isfile(REF + "/optim/fractional.py")


def load_reference_optim():
  """-> (fractional, visibility_aware) modules of the REAL reference (pure torch), with its two Taichi kernel modules
  replaced by the restated kernels above."""
  def _ours(k):
    return k == "taichi_splatting" or k.startswith("taichi_splatting.")
  saved = {k: v for k, v in sys.modules.items() if _ours(k)}
  try:
    pkg = types.ModuleType("taichi_splatting"); pkg.__path__ = []
    opt = types.ModuleType("taichi_splatting.optim"); opt.__path__ = [REF + "/optim"]
    sys.modules["taichi_splatting"], sys.modules["taichi_splatting.optim"] = pkg, opt
    for name, alg in (("fractional_adam", ADAM), ("fractional_laprop", LAPROP)):
      m = kernel_module(alg)
      sys.modules["taichi_splatting.optim." + name] = m
      setattr(opt, name, m)

    def _load(name):
      spec = importlib.util.spec_from_file_location("taichi_splatting.optim." + name, f"{REF}/optim/{name}.py")
      m = importlib.util.module_from_spec(spec)
      sys.modules["taichi_splatting.optim." + name] = m
      spec.loader.exec_module(m)
      setattr(opt, name, m)
      return m
    _load("util")
    return _load("fractional"), _load("visibility_aware")
  finally:
    for k in [k for k in sys.modules if _ours(k)]:
      del sys.modules[k]
    sys.modules.update(saved)
