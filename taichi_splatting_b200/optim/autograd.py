"""restore_grad: a scope in which the given tensors carry fresh zero gradients; whatever gradients they had before
come back on exit, also when the body raises (reference behaviour: taichi_splatting/optim/autograd.py:5-16)."""
import contextlib

import torch


class restore_grad(contextlib.ContextDecorator):
  def __init__(self, *tensors):
    self._tensors = tensors
    self._previous = None

  def __enter__(self):
    self._previous = tuple(t.grad for t in self._tensors)
    for t in self._tensors:
      if t.requires_grad is True:
        t.grad = torch.zeros_like(t)
    return self

  def __exit__(self, exc_type, exc, tb):
    for t, g in zip(self._tensors, self._previous):
      t.grad = g
    return False
