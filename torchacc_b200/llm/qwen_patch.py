"""Qwen (remote-code, v1) attention patch (reference torchacc/llm/qwen_patch.py:9-53).

The reference rewrites the *source text* of the model's attention ``forward`` with regexes to drop dtype/cuda
asserts and to swap ``flash_attn_unpadded_func`` for its XLA op.  Here the same effect is achieved by rebinding
the module-level symbols the remote code looks up at call time -- no source rewriting, so it survives upstream
formatting changes."""
from __future__ import annotations

import sys
import types

import torch

from ..utils.logger import logger


def _unpadded_func(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, dropout_p=0.0, softmax_scale=None,
                   causal=False, **kwargs):
    """Signature of flash-attn v1/v2 ``flash_attn_unpadded_func`` / ``flash_attn_varlen_func`` on packed tokens."""
    from ..ops.attention import flash_attn_varlen_cu
    import math
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(q.shape[-1])
    return flash_attn_varlen_cu(q, k, v, cu_seqlens_q.int(), cu_seqlens_k.int(), scale, causal, (-1, -1))


def rewrite_forward(module: types.ModuleType) -> bool:
    """Rebind flash-attention entry points inside a remote-code modeling module."""
    done = False
    for name in ("flash_attn_unpadded_func", "flash_attn_varlen_func"):
        if hasattr(module, name):
            setattr(module, name, _unpadded_func)
            done = True
    return done


def patch_qwen_model(model: torch.nn.Module) -> torch.nn.Module:
    """Patch the modeling module that defines ``model`` (and enable its flash-attention path)."""
    mod = sys.modules.get(type(model).__module__)
    if mod is None or not rewrite_forward(mod):
        logger.warning("patch_qwen_model: no flash-attention entry point found in %s", type(model).__module__)
    for m in model.modules():
        if hasattr(m, "use_flash_attn"):
            m.use_flash_attn = True
    return model
