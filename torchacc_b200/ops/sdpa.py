"""``scaled_dot_product_attention`` replacement (reference torchacc/ops/scaled_dot_product_attention.py:7-20).

The reference shim treats ANY ``attn_mask`` as "causal" and ignores ``scale``.  Here:
``is_causal`` and ``scale`` are honoured; a boolean/additive ``attn_mask`` falls back to PyTorch's SDPA because
the flash kernels only express causal / sliding-window / per-sequence length masks."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .attention import flash_attn_func

_ORIG = F.scaled_dot_product_attention


def scaled_dot_product_attention(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None,
                                 enable_gqa=False, **kwargs):
    """query/key/value: [B, H, S, D] (PyTorch layout)."""
    if attn_mask is not None or query.dim() != 4 or query.dtype not in (torch.bfloat16, torch.float16):
        return _ORIG(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale,
                     **({"enable_gqa": enable_gqa} if enable_gqa else {}))
    q, k, v = query.transpose(1, 2), key.transpose(1, 2), value.transpose(1, 2)
    out = flash_attn_func(q, k, v, dropout_p=dropout_p, softmax_scale=scale, causal=is_causal)
    return out.transpose(1, 2)


def patch_sdpa():
    F.scaled_dot_product_attention = scaled_dot_product_attention
    torch.nn.functional.scaled_dot_product_attention = scaled_dot_product_attention


def unpatch_sdpa():
    F.scaled_dot_product_attention = _ORIG
