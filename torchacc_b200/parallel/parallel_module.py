"""Base class of the parallel wrappers (reference torchacc/dist/parallel_module.py:8-69)."""
from __future__ import annotations

import torch
import torch.nn as nn


class ParallelModule(nn.Module):
    """Holds the wrapped model, the config, the device and the mesh; subclasses add one parallel strategy."""

    def __init__(self, model: nn.Module, config, **kwargs):
        super().__init__()
        self.model = model
        self._config = config
        from .bootstrap import current_device
        self.device = current_device()
        self.mesh = config.get_mesh()
        d = config.dist
        self.has_dp = (d.dp.size or 1) > 1
        self.has_tp = d.tp.size > 1
        self.has_pp = d.pp.size > 1
        self.has_fsdp = d.fsdp.size > 1
        self.has_sp = d.sp.size > 1
        self.spmd_fsdp = False

    def _get_underlay_model(self):
        return self.model

    def _update_underlay_model(self, model):
        self.model = model

    def clip_grad_norm_(self, max_norm, norm_type=2.0):
        params = [p for p in self.parameters() if p.grad is not None]
        return torch.nn.utils.clip_grad_norm_(params, max_norm, norm_type)

    def forward_backward(self, *args, output_fn=None, **kwargs):
        """Only meaningful with pipeline parallelism (reference parallel_module.py:52-69)."""
        raise NotImplementedError("forward_backward is provided by PipelineParallel")

    def forward(self, *args, **kwargs):
        return self.model(*args, **kwargs)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.__dict__["_modules"]["model"], name)
