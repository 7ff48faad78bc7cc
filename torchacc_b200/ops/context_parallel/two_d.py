"""2-D context parallelism ("FlashSequence"; reference torchacc/ops/context_parallel/context_parallel_2d.py:11-127):
Ulysses all-to-all inside the intra group x ring attention across the inter group."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .comm import _world, seq_head_all_to_all
from .ring import ring_attention
from .ulysses import ulysses


def context_parallel_2d(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, q_lens: Optional[torch.Tensor] = None,
                        k_lens: Optional[torch.Tensor] = None, dropout_p: float = 0.0,
                        softmax_scale: Optional[float] = None, causal: bool = False, window_size: tuple = (-1, -1),
                        alibi_slopes: Optional[tuple] = None, deterministic: bool = False,
                        inter_process_group: Optional[dist.ProcessGroup] = None,
                        intra_process_group: Optional[dist.ProcessGroup] = None, zigzag: bool = False) -> torch.Tensor:
    """q: [B, S/cp, Hq, D] where cp = inter x intra and the sequence is laid out inter-major (rank = inter*intra_size
    + intra holds chunk ``rank``)."""
    n_intra, n_inter = _world(intra_process_group), _world(inter_process_group)
    if n_inter == 1:
        return ulysses(q, k, v, q_lens, k_lens, dropout_p, softmax_scale, causal, window_size, alibi_slopes,
                       deterministic, intra_process_group)
    if n_intra == 1:
        return ring_attention(q, k, v, q_lens, k_lens, dropout_p, softmax_scale, causal, window_size, alibi_slopes,
                              deterministic, inter_process_group, zigzag=zigzag)
    # heads scattered / sequence gathered inside the intra group ...
    q2 = seq_head_all_to_all(q, 2, 1, intra_process_group)
    k2 = seq_head_all_to_all(k, 2, 1, intra_process_group)
    v2 = seq_head_all_to_all(v, 2, 1, intra_process_group)
    # ... ring over the inter group on [B, S/inter, H/intra, D]
    out = ring_attention(q2, k2, v2, None, k_lens, dropout_p, softmax_scale, causal, window_size, alibi_slopes,
                         deterministic, inter_process_group, zigzag=zigzag)
    return seq_head_all_to_all(out, 1, 2, intra_process_group)
