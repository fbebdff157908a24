"""Sparse, visibility-weighted optimisers over the visible set of a render (the reference's `optim` package minus
`ParameterClass`, its tensordict container, which is not on the hot path)."""
from . import autograd, fractional, visibility_aware
from .autograd import restore_grad
from .fractional import ADAM, LAPROP, FractionalAdam, FractionalLaProp, FractionalOpt, SparseAdam, SparseLaProp
from .visibility_aware import VisibilityAwareAdam, VisibilityAwareLaProp, VisibilityOptimizer

__all__ = [name for name, obj in list(globals().items())
           if not name.startswith("_") and (isinstance(obj, type) or name in ("ADAM", "LAPROP"))]
