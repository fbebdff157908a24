"""restore_grad (reference: taichi_splatting/optim/autograd.py:5-16): run a block with fresh zero gradients and put
the previous ones back afterwards."""
from contextlib import contextmanager

import torch


@contextmanager
def restore_grad(*tensors):
  saved = [t.grad for t in tensors]
  try:
    for t in tensors:
      if t.requires_grad is True:
        t.grad = torch.zeros_like(t)
    yield
  finally:
    for t, g in zip(tensors, saved):
      t.grad = g
