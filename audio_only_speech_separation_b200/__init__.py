"""B200-native (sm_100a) implementation of the look2hear dual-path separation hot path.

``models`` and ``losses`` mirror ``look2hear.models`` / ``look2hear.losses`` of the reference for the path
(``TasNet`` with ``module="DPRNN"`` / ``"DPTNet"``, ``Sepformer``, ``PITLossWrapper``, ``pairwise_neg_*``) so that the reference's
``getattr(look2hear.models, name)`` / ``getattr(look2hear.losses, name)`` lookups resolve unchanged.
All arithmetic runs in the hand-written CUDA library ``libdualpath_b200.so`` (C-ABI: ``include/dualpath_b200.h``);
``torch_ops`` registers the operator-level entry points as ``torch.library`` custom ops (``torch.ops.dualpath.*``).
"""
from . import _lib  # noqa: F401
from . import torch_ops  # noqa: F401  (registers torch.ops.dualpath.*; CUDA dispatch key only)

__version__ = "0.1.0"
