"""Micro-batch splitting of call arguments (reference torchacc/dist/pp/microbatch.py:7-48)."""
from __future__ import annotations

import inspect
from typing import Any, Dict, List, Tuple

import torch


def bind_args_to_kwargs(args: tuple, kwargs: dict, signature: inspect.Signature) -> Dict[str, Any]:
    """Positional + keyword call arguments -> one name->value dict following ``signature``."""
    names = [n for n in signature.parameters if n != "self"]
    if len(args) > len(names):
        raise TypeError(f"too many positional arguments ({len(args)} > {len(names)})")
    bound = dict(zip(names, args))
    dup = set(bound) & set(kwargs)
    if dup:
        raise TypeError(f"multiple values for argument(s) {sorted(dup)}")
    bound.update(kwargs)
    return bound


def split_kwargs_into_chunks(kwargs: Dict[str, Any], chunks: int) -> List[Dict[str, Any]]:
    """Chunk every tensor along dim 0; non-tensors (None, ints, bools...) are replicated."""
    out = [dict() for _ in range(chunks)]
    for k, v in kwargs.items():
        if isinstance(v, torch.Tensor) and v.dim() > 0:
            if v.shape[0] % chunks != 0:
                raise ValueError(f"batch dimension of '{k}' ({v.shape[0]}) is not divisible by {chunks} micro-batches")
            for i, piece in enumerate(v.chunk(chunks, dim=0)):
                out[i][k] = piece
        else:
            for i in range(chunks):
                out[i][k] = v
    return out


def split_args_kwargs_into_chunks(args: tuple, kwargs: dict, chunks: int) -> Tuple[List[tuple], List[dict]]:
    a = split_kwargs_into_chunks({i: v for i, v in enumerate(args)}, chunks)
    k = split_kwargs_into_chunks(kwargs, chunks)
    return [tuple(d[i] for i in range(len(args))) for d in a], k
