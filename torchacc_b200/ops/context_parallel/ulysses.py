"""DeepSpeed-Ulysses attention (reference torchacc/ops/context_parallel/ulysses.py:9-77).

Each rank holds a sequence shard ``[B, S/cp, H, D]``.  One all-to-all turns q/k/v into ``[B, S, H/cp, D]`` (full
sequence, a slice of the heads), attention runs locally on our flash kernels, and one all-to-all brings the output
back to sequence shards.  q, k and v travel in ONE packed exchange when the head counts allow it.  ``rope_func`` is
applied after the exchange, when every rank sees whole sequences (same hook as the reference, :59-60)."""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from ..attention import flash_attn_func, flash_attn_varlen_func
from .comm import _world, seq_head_all_to_all


def _lens_to_mask(lens: torch.Tensor, S: int) -> torch.Tensor:
    return (torch.arange(S, device=lens.device)[None, :] < lens[:, None]).to(torch.int32)


def ulysses(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, q_lens: Optional[torch.Tensor] = None,
            k_lens: Optional[torch.Tensor] = None, dropout_p: float = 0.0, softmax_scale: Optional[float] = None,
            causal: bool = False, window_size: tuple = (-1, -1), alibi_slopes: Optional[tuple] = None,
            deterministic: bool = False, process_group: Optional[dist.ProcessGroup] = None, position_ids=None,
            rope_func: Callable = None) -> torch.Tensor:
    """q: [B, S/cp, Hq, D]; k, v: [B, S/cp, Hk, D] -> [B, S/cp, Hq, D].  ``q_lens``/``k_lens`` are the true (global)
    sequence lengths per batch entry for padded batches."""
    cp = _world(process_group)
    Hq, Hk = q.shape[2], k.shape[2]
    if Hq % cp != 0 or Hk % cp != 0:
        raise ValueError(f"Ulysses needs the head counts ({Hq}, {Hk}) to be divisible by the group size {cp}")
    if cp > 1:
        # one packed exchange: [B, S/cp, Hq + 2 Hk, D] is not head-sliceable per tensor, so interleave per rank slice
        gq, gk = Hq // cp, Hk // cp
        packed = torch.cat([q.reshape(*q.shape[:2], cp, gq, -1), k.reshape(*k.shape[:2], cp, gk, -1),
                            v.reshape(*v.shape[:2], cp, gk, -1)], dim=3)          # [B, S/cp, cp, gq+2gk, D]
        packed = packed.reshape(*q.shape[:2], cp * (gq + 2 * gk), q.shape[-1])
        full = seq_head_all_to_all(packed, 2, 1, process_group)                    # [B, S, gq+2gk, D]
        q, k, v = full.split([gq, gk, gk], dim=2)
    if rope_func is not None:
        q, k, v = rope_func(q, k, v) if position_ids is None else rope_func(q, k, v, position_ids)
    if q_lens is not None or k_lens is not None:
        lens = k_lens if k_lens is not None else q_lens
        out = flash_attn_varlen_func(q.contiguous(), k.contiguous(), v.contiguous(), _lens_to_mask(lens, q.shape[1]),
                                     dropout_p, softmax_scale, causal, window_size, alibi_slopes, deterministic)
    else:
        out = flash_attn_func(q.contiguous(), k.contiguous(), v.contiguous(), dropout_p, softmax_scale, causal,
                              window_size, alibi_slopes, deterministic)
    if cp > 1:
        out = seq_head_all_to_all(out, 1, 2, process_group)                       # back to [B, S/cp, Hq, D]
    return out
