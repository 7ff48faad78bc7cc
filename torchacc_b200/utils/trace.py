"""Symbolic tracing helpers for pipeline-stage construction (reference torchacc/utils/trace.py:21-175).

Pipeline parallelism in this framework does NOT require an fx trace: stages are cut on the module tree
(parallel/pp/partition.py).  ``trace`` is still provided for models a user wants as a ``GraphModule`` -- HF models
go through ``transformers.utils.fx`` when available, everything else through ``torch.fx``."""
from __future__ import annotations

import inspect
from typing import List, Optional

import torch
import torch.fx as fx


def get_concrete_args(model: torch.nn.Module, input_names: List[str]) -> dict:
    """Arguments of ``model.forward`` not listed in ``input_names`` are frozen to their defaults."""
    sig = inspect.signature(model.forward)
    return {p.name: p.default for p in sig.parameters.values()
            if p.name not in input_names and p.default is not inspect.Parameter.empty}


def trace(model: torch.nn.Module, input_names: Optional[List[str]] = None) -> fx.GraphModule:
    input_names = input_names or list(inspect.signature(model.forward).parameters)[:1]
    try:
        from transformers import PreTrainedModel
        if isinstance(model, PreTrainedModel):
            from transformers.utils.fx import symbolic_trace
            return symbolic_trace(model, input_names=input_names)
    except Exception:
        pass
    tracer = fx.Tracer()
    graph = tracer.trace(model, concrete_args=get_concrete_args(model, input_names))
    return fx.GraphModule(model, graph)
