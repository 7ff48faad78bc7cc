"""Context parallelism for HuggingFace causal LMs handed to ``accelerate()`` (``dist.sp.size > 1``).

The reference ships context parallelism as attention functions only (torchacc/ops/context_parallel/ulysses.py:9-77,
ring_attn.py:275-330, context_parallel_2d.py:12-120); a user has to call them from their own attention module and
split the batch by hand.  Here the two halves are done for any HF model that dispatches through the attention-interface
registry (Llama, Qwen2/3, Mistral, Gemma, ...):

* :class:`HFContextParallel` shards ``input_ids`` / ``labels`` / ``position_ids`` along the sequence, pre-shifts the
  labels globally (so the token at a shard boundary still predicts the first token of the next shard) and hands them
  over as ``shift_labels``;
* the model's attention implementation becomes ``"torchacc_b200_cp"``: q/k/v arrive RoPE-rotated with GLOBAL
  positions and go through Ulysses / ring / 2-D attention over the sp group.

The sharding engine averages gradients over the sp ranks (they are replicas of each parameter shard), so the local
mean loss is weighted by ``local valid labels x cp / global valid labels``: the gradient is that of the global token
mean, and the returned loss VALUE is the global mean on every rank.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from .comm import _rank, _world
from .ring import ring_attention, zigzag_split
from .two_d import context_parallel_2d
from .ulysses import ulysses

IMPL_NAME = "torchacc_b200_cp"


class _CPContext:
    __slots__ = ("mesh", "mode", "zigzag")

    def __init__(self, mesh, mode: str):
        self.mesh = mesh
        self.mode = mode or "ulysses"
        self.zigzag = self.mode == "ring" and getattr(mesh, "zigzag", True)

    @property
    def group(self):
        return self.mesh.get_sp_proc_group()

    def shard(self, x: Optional[torch.Tensor], dim: int = 1) -> Optional[torch.Tensor]:
        if x is None:
            return None
        g = self.group
        if self.zigzag:
            return zigzag_split(x, dim, g)
        return x.chunk(_world(g), dim=dim)[_rank(g)].contiguous()


def _cp_attention_interface(module, query, key, value, attention_mask, dropout=0.0, scaling=None,
                            sliding_window=None, **kwargs):
    """HF attention-interface signature: q [B, Hq, S_local, D], k/v [B, Hk, S_local, D] -> ([B, S_local, Hq, D], None)."""
    ctx: Optional[_CPContext] = getattr(module, "_tb_cp", None)
    q, k, v = query.transpose(1, 2), key.transpose(1, 2), value.transpose(1, 2)
    causal = bool(getattr(module, "is_causal", True))
    window = (sliding_window - 1, 0) if sliding_window else (-1, -1)
    if ctx is None:
        from ..attention import flash_attn_func
        return flash_attn_func(q.contiguous(), k.contiguous(), v.contiguous(), dropout, scaling, causal, window), None
    if attention_mask is not None:
        raise ValueError("context parallelism takes un-padded batches (attention_mask must be None)")
    if ctx.mode == "ulysses":
        out = ulysses(q, k, v, dropout_p=dropout, softmax_scale=scaling, causal=causal, window_size=window,
                      process_group=ctx.group)
    elif ctx.mode == "ring":
        out = ring_attention(q, k, v, softmax_scale=scaling, causal=causal, window_size=window,
                             process_group=ctx.group, zigzag=ctx.zigzag)
    else:
        out = context_parallel_2d(q, k, v, softmax_scale=scaling, causal=causal, window_size=window,
                                  inter_process_group=ctx.mesh.get_ring_proc_group(),
                                  intra_process_group=ctx.mesh.get_ulysses_proc_group(), zigzag=False)
    return out, None


def _register() -> None:
    from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
    ALL_ATTENTION_FUNCTIONS[IMPL_NAME] = _cp_attention_interface


def is_hf_causal_lm(model: nn.Module) -> bool:
    cfg = getattr(model, "config", None)
    return cfg is not None and hasattr(cfg, "_attn_implementation") and hasattr(model, "loss_function")


class HFContextParallel(nn.Module):
    """Wraps an HF ``*ForCausalLM``: sequence sharding on the way in, context-parallel attention inside."""

    def __init__(self, model: nn.Module, mesh, mode: str = "ulysses"):
        super().__init__()
        _register()
        self.model = model
        self.ctx = _CPContext(mesh, mode)
        model.config._attn_implementation = IMPL_NAME
        n = 0
        for m in model.modules():
            if hasattr(m, "q_proj") or hasattr(m, "qkv_proj") or type(m).__name__.endswith("Attention"):
                m._tb_cp = self.ctx
                n += 1
        if n == 0:
            raise ValueError(f"{type(model).__name__}: no attention modules found for context parallelism")

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.__dict__["_modules"]["model"], name)

    def forward(self, input_ids=None, labels=None, position_ids=None, attention_mask=None, inputs_embeds=None,
                **kwargs):
        ctx = self.ctx
        full = input_ids if input_ids is not None else inputs_embeds
        B, S = full.shape[:2]
        cp = _world(ctx.group)
        if S % (2 * cp if ctx.zigzag else cp) != 0:
            raise ValueError(f"sequence length {S} is not divisible by the context-parallel layout ({cp} ranks)")
        if attention_mask is not None and not bool(attention_mask.all()):
            raise ValueError("context parallelism takes un-padded (packed) batches")
        if position_ids is None:
            position_ids = torch.arange(S, device=full.device).unsqueeze(0).expand(B, S)
        extra = {}
        n_local = n_total = None
        if labels is not None:
            shift = kwargs.pop("shift_labels", None)
            if shift is None:
                shift = torch.nn.functional.pad(labels[:, 1:], (0, 1), value=-100)
            shift_local = ctx.shard(shift)
            extra["labels"] = ctx.shard(labels)
            extra["shift_labels"] = shift_local
            n_local = (shift_local != -100).sum()
        out = self.model(input_ids=ctx.shard(input_ids), inputs_embeds=ctx.shard(inputs_embeds),
                         position_ids=ctx.shard(position_ids), attention_mask=None, **extra, **kwargs)
        loss = getattr(out, "loss", None)
        if loss is not None and cp > 1:
            # exact global mean: the engine AVERAGES gradients over the sp replicas, so the local mean is weighted by
            # (local valid labels x cp / global valid labels); the reported value is the global mean on every rank
            with torch.no_grad():
                stat = torch.stack([loss.detach().float() * n_local, n_local.float()])
                dist.all_reduce(stat, group=ctx.group)
                n_total = stat[1].clamp(min=1.0)
                global_loss = stat[0] / n_total
                weight = n_local.float() * cp / n_total
            weighted = loss * weight.to(loss.dtype)
            out.loss = weighted + (global_loss - weighted.detach()).to(loss.dtype)
        return out
