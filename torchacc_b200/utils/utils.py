"""Small shared helpers (reference torchacc/utils/utils.py:15-373: partitioners, class lookup, fp32 output
conversion, pytree map)."""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, List, Optional, Sequence

import torch
import torch.nn as nn


def apply_to_tensors(fn: Callable[[torch.Tensor], Any], obj: Any) -> Any:
    """Map ``fn`` over every tensor in a nested container, preserving container types (incl. dict subclasses,
    namedtuples, dataclasses and HF ``ModelOutput``)."""
    if isinstance(obj, torch.Tensor):
        return fn(obj)
    if isinstance(obj, dict):
        try:
            out = type(obj)()
            for k, v in obj.items():
                out[k] = apply_to_tensors(fn, v)
            return out
        except Exception:
            return {k: apply_to_tensors(fn, v) for k, v in obj.items()}
    if isinstance(obj, tuple) and hasattr(obj, "_fields"):
        return type(obj)(*(apply_to_tensors(fn, v) for v in obj))
    if isinstance(obj, (list, tuple, set)):
        return type(obj)(apply_to_tensors(fn, v) for v in obj)
    if dataclasses.is_dataclass(obj) and not isinstance(obj, type):
        return dataclasses.replace(obj, **{f.name: apply_to_tensors(fn, getattr(obj, f.name))
                                           for f in dataclasses.fields(obj)})
    return obj


def convert_to_fp32(obj: Any) -> Any:
    """Cast floating-point (bf16/fp16) tensors in ``obj`` to fp32 (reference utils.py:281-339)."""
    def cast(t):
        return t.float() if t.is_floating_point() and t.dtype in (torch.float16, torch.bfloat16) else t
    return apply_to_tensors(cast, obj)


class ConvertOutputsToFp32:
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, *a, **k):
        return convert_to_fp32(self.fn(*a, **k))


convert_outputs_to_fp32 = ConvertOutputsToFp32


def get_module_class_from_name(module: nn.Module, name: str) -> Optional[type]:
    for m in module.modules():
        if type(m).__name__ == name:
            return type(m)
    return None


def call_to_str(base: str, *args, **kwargs) -> str:
    parts = [repr(a) for a in args] + [f"{k}={v!r}" for k, v in kwargs.items()]
    return f"{base}({', '.join(parts)})"


def partition_uniform(num_items: int, num_parts: int) -> List[int]:
    """Boundaries of an even split: ``parts[i]..parts[i+1]`` is part i."""
    base, rem = divmod(num_items, num_parts)
    bounds = [0]
    for i in range(num_parts):
        bounds.append(bounds[-1] + base + (1 if i < rem else 0))
    return bounds


def partition_balanced(weights: Sequence[float], num_parts: int) -> List[int]:
    """Contiguous partition minimising the heaviest part (binary search on the bottleneck + greedy fill)."""
    n = len(weights)
    if num_parts >= n:
        return list(range(n + 1)) + [n] * (num_parts - n)
    lo, hi = max(weights), sum(weights)

    def parts_needed(cap):
        cnt, cur = 1, 0.0
        for w in weights:
            if cur + w > cap:
                cnt, cur = cnt + 1, w
            else:
                cur += w
        return cnt

    for _ in range(64):
        mid = (lo + hi) / 2
        if parts_needed(mid) <= num_parts:
            hi = mid
        else:
            lo = mid
    bounds, cur = [0], 0.0
    for i, w in enumerate(weights):
        remaining_items, remaining_parts = n - i, num_parts - len(bounds) + 1
        if (cur + w > hi * (1 + 1e-9) and cur > 0) or remaining_items < remaining_parts:
            if len(bounds) < num_parts:
                bounds.append(i)
                cur = 0.0
        cur += w
    while len(bounds) < num_parts:
        bounds.append(n)
    bounds.append(n)
    return bounds
