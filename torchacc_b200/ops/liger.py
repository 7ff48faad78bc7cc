"""Kernel patches for HuggingFace models (reference torchacc/ops/liger.py:10-153).

The reference swaps HF module forwards onto liger's Triton kernels.  Same entry points here
(``apply_liger_kernel``, ``apply_liger_kernel_to_llama``, ``apply_liger_kernel_to_qwen2``) but the replacements are
our sm_100a kernels: RMSNorm, SwiGLU, RoPE-free fused loss (linear + cross-entropy), and the tcgen05 GEMM behind
every ``nn.Linear`` of the model.  Patches are applied to *classes* (like liger) or to one model instance.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ..utils.logger import logger
from .cross_entropy import fused_linear_cross_entropy
from .linear import linear
from .rmsnorm import rmsnorm
from .swiglu import swiglu_separate


def rms_forward(self, hidden_states):
    eps = getattr(self, "variance_epsilon", getattr(self, "eps", 1e-6))
    y, _ = rmsnorm(hidden_states, self.weight, eps)
    return y


def mlp_forward(self, x):
    g = linear(x, self.gate_proj.weight, self.gate_proj.bias)
    u = linear(x, self.up_proj.weight, self.up_proj.bias)
    return linear(swiglu_separate(g, u), self.down_proj.weight, self.down_proj.bias)


def linear_forward(self, x):
    return linear(x, self.weight, self.bias)


def _make_lce_forward(orig_forward):
    """Causal-LM forward that computes the loss with fused linear+CE when labels are given."""

    def lce_forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                    inputs_embeds=None, labels=None, use_cache=None, **kwargs):
        if labels is None:
            return orig_forward(self, input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                                past_key_values=past_key_values, inputs_embeds=inputs_embeds, use_cache=use_cache,
                                **kwargs)
        outputs = self.model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                             past_key_values=past_key_values, inputs_embeds=inputs_embeds, use_cache=use_cache,
                             **{k: v for k, v in kwargs.items() if k not in ("num_items_in_batch", "logits_to_keep")})
        hidden = outputs[0]
        shift = torch.full_like(labels, -100)
        shift[..., :-1] = labels[..., 1:]
        loss = fused_linear_cross_entropy(hidden.reshape(-1, hidden.shape[-1]), self.lm_head.weight,
                                          shift.reshape(-1))
        from transformers.modeling_outputs import CausalLMOutputWithPast
        return CausalLMOutputWithPast(loss=loss, logits=None, past_key_values=getattr(outputs, "past_key_values", None))

    return lce_forward


def _patch_family(modeling, prefix: str, rms_norm: bool, swiglu: bool, fused_linear_cross_entropy_: bool):
    if rms_norm and hasattr(modeling, f"{prefix}RMSNorm"):
        getattr(modeling, f"{prefix}RMSNorm").forward = rms_forward
    if swiglu and hasattr(modeling, f"{prefix}MLP"):
        getattr(modeling, f"{prefix}MLP").forward = mlp_forward
    if fused_linear_cross_entropy_ and hasattr(modeling, f"{prefix}ForCausalLM"):
        cls = getattr(modeling, f"{prefix}ForCausalLM")
        if not getattr(cls, "_tb_lce_patched", False):
            cls.forward = _make_lce_forward(cls.forward)
            cls._tb_lce_patched = True


def apply_liger_kernel_to_llama(rope: bool = True, cross_entropy: bool = False, fused_linear_cross_entropy: bool = True,
                                rms_norm: bool = True, swiglu: bool = True, model: Optional[nn.Module] = None) -> None:
    """Reference liger.py:32-82.  ``rope``: HF applies RoPE inside its attention in a layout our in-place kernel
    does not see; attention itself is redirected by ``utils.patch.patch_fa`` so the flag is accepted and ignored."""
    from transformers.models.llama import modeling_llama
    _patch_family(modeling_llama, "Llama", rms_norm, swiglu, fused_linear_cross_entropy)


def apply_liger_kernel_to_qwen2(rope: bool = True, cross_entropy: bool = False, fused_linear_cross_entropy: bool = True,
                                rms_norm: bool = True, swiglu: bool = True, model: Optional[nn.Module] = None) -> None:
    """Reference liger.py:86-130."""
    from transformers.models.qwen2 import modeling_qwen2
    _patch_family(modeling_qwen2, "Qwen2", rms_norm, swiglu, fused_linear_cross_entropy)


def patch_linears(model: nn.Module) -> int:
    """Route every ``nn.Linear`` of this model instance through the tcgen05 GEMM."""
    n = 0
    for m in model.modules():
        if type(m) is nn.Linear:
            m.forward = linear_forward.__get__(m, nn.Linear)
            n += 1
    return n


def apply_liger_kernel(model: Optional[nn.Module] = None) -> None:
    """Best-effort patching used by ``accelerate()`` (reference liger.py:133-153).  With a model, only the HF
    families that actually occur in it are patched (class-level patches are process-global; a native model must not
    change how unrelated HF models behave); without one, every supported family is patched like the reference."""
    try:
        import transformers  # noqa: F401
    except Exception:
        transformers = None
    if transformers is not None:
        families = {"transformers.models.llama.": apply_liger_kernel_to_llama,
                    "transformers.models.qwen2.": apply_liger_kernel_to_qwen2}
        if model is None:
            wanted = list(families.values())
        else:
            mods = {type(m).__module__ for m in model.modules()}
            wanted = [fn for prefix, fn in families.items() if any(x.startswith(prefix) for x in mods)]
        for fn in wanted:
            try:
                fn()
            except Exception as e:  # pragma: no cover - depends on the installed transformers
                logger.debug("kernel patch %s skipped: %s", fn.__name__, e)
    if model is not None and not type(model).__module__.startswith("torchacc_b200"):
        patch_linears(model)
