"""Gradient (activation) checkpointing.

Reference: torchacc/utils/checkpoint.py:14-81 (``checkpoint_module``, ``gradient_checkpoint``, ``fx_checkpoint``);
there both names only exist when torch_xla is importable, so the eager backend trains without GC (SURVEY
Appendix B #1).  Here checkpointing is a thin wrapper module around ``torch.utils.checkpoint`` (non-reentrant), and
inside FSDP units it is applied *inside* the unit so recomputation reuses the gathered parameters.
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence, Union

import torch
import torch.nn as nn
from torch.utils.checkpoint import checkpoint as _ckpt


class CheckpointedModule(nn.Module):
    """Runs the wrapped module under activation checkpointing when gradients are enabled."""

    def __init__(self, module: nn.Module):
        super().__init__()
        self.module = module

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.__dict__["_modules"]["module"], name)

    def forward(self, *args, **kwargs):
        if torch.is_grad_enabled():
            from ..ops.swiglu import activations_recomputed_later
            with activations_recomputed_later():     # what this pass saves is dropped: fused ops skip writing it
                return _ckpt(self.module, *args, use_reentrant=False, **kwargs)
        return self.module(*args, **kwargs)


def checkpoint_module(module: nn.Module) -> nn.Module:
    """Wrap one module (reference checkpoint.py:14)."""
    return module if isinstance(module, CheckpointedModule) else CheckpointedModule(module)


def _resolve(model: nn.Module, gc_cls: Iterable[Union[str, type]]):
    classes = []
    names = set()
    for c in gc_cls:
        (classes if isinstance(c, type) else names).__iadd__([c]) if isinstance(c, type) else names.add(c)
    for m in model.modules():
        if type(m).__name__ in names and type(m) not in classes:
            classes.append(type(m))
    return tuple(classes)


def gradient_checkpoint(model: nn.Module, gc_cls: Iterable[Union[str, type]], gc_cnt: Optional[int] = None,
                        skip_types: Sequence[type] = ()) -> nn.Module:
    """Checkpoint every submodule whose class (or class name) is in ``gc_cls`` (reference checkpoint.py:67-81).
    ``gc_cnt`` limits the number of wrapped instances (first N in module order)."""
    classes = _resolve(model, gc_cls)
    if not classes:
        return model
    left = [gc_cnt if gc_cnt is not None else float("inf")]
    wrapped = []

    def recurse(parent):
        for name, child in list(parent.named_children()):
            if isinstance(child, (CheckpointedModule, *skip_types)):
                continue
            if isinstance(child, classes) and left[0] > 0:
                parent._modules[name] = CheckpointedModule(child)
                left[0] -= 1
                wrapped.append(name)
            else:
                recurse(child)

    recurse(model)
    if wrapped:
        disable_kv_cache(model)
    return model


def disable_kv_cache(model: nn.Module) -> None:
    """HF models: a KV cache filled in the forward would be appended to again by the recomputation, so training
    with activation checkpointing turns ``config.use_cache`` off (what HF's own gradient_checkpointing does)."""
    for m in model.modules():
        cfg = getattr(m, "config", None)
        if cfg is not None and getattr(cfg, "use_cache", False):
            try:
                cfg.use_cache = False
            except Exception:
                pass


def fx_checkpoint(graph_module, gc_cls):
    """API-compat shim for the reference's fx-graph variant (checkpoint.py:17-64): our pipeline stages keep real
    submodules, so class-based wrapping is sufficient."""
    return gradient_checkpoint(graph_module, gc_cls)
