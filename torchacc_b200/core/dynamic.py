"""``mark_dynamic`` (reference torchacc/core/dynamic.py:13-46 marks bounded-dynamic dims for the XLA compiler).
Eager CUDA kernels take runtime shapes, so this only validates its arguments and records the bounds on the tensor
(``AsyncLoader`` bucketing remains the mechanism that keeps shapes in a small set)."""
from __future__ import annotations

from typing import Sequence, Union

import torch


def mark_dynamic(x: torch.Tensor, dims: Union[int, Sequence[int]], bounds: Union[int, Sequence[int]]) -> torch.Tensor:
    dims = [dims] if isinstance(dims, int) else list(dims)
    bounds = [bounds] if isinstance(bounds, int) else list(bounds)
    if len(dims) != len(bounds):
        raise ValueError("dims and bounds must have the same length")
    for d, b in zip(dims, bounds):
        if not -x.dim() <= d < x.dim():
            raise ValueError(f"dim {d} out of range for a {x.dim()}-d tensor")
        if x.shape[d] > b:
            raise ValueError(f"dim {d} has size {x.shape[d]} > bound {b}")
    x._tb_dynamic_bounds = dict(zip(dims, bounds))
    return x
