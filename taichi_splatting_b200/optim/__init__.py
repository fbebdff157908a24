"""Sparse, visibility-weighted optimisers (reference: taichi_splatting/optim/__init__.py).

`ParameterClass` (a tensordict container around these optimisers, optim/parameter_class.py) is not part of the
hot path and is not provided."""
from .autograd import restore_grad
from .fractional import FractionalAdam, FractionalLaProp, FractionalOpt, SparseAdam, SparseLaProp
from .visibility_aware import VisibilityAwareAdam, VisibilityAwareLaProp, VisibilityOptimizer

__all__ = ['FractionalOpt', 'FractionalAdam', 'FractionalLaProp', 'SparseAdam', 'SparseLaProp',
           'VisibilityAwareAdam', 'VisibilityAwareLaProp', 'VisibilityOptimizer', 'restore_grad']
