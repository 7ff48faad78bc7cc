"""``PipelineParallel`` wrapper (reference torchacc/dist/pp/pipeline.py:29-149)."""
from __future__ import annotations

import inspect

import torch
import torch.nn as nn

from ...utils.logger import log_rank0
from ..parallel_module import ParallelModule
from .executor import PipeExecutor
from .microbatch import bind_args_to_kwargs
from .partition import build_stages, names_of_split_points


def preprocess_config(config, model: nn.Module):
    """Module objects in ``pp.split_points`` -> qualified names (reference pipeline.py:13-25)."""
    config.dist.pp.split_points = names_of_split_points(model, config.dist.pp.split_points)
    return config


class PipelineParallel(ParallelModule):

    def __init__(self, model: nn.Module, config, orig_forward_sig=None, **kwargs):
        super().__init__(model, config)
        pp = config.dist.pp
        self.orig_forward_sig = orig_forward_sig or inspect.signature(model.forward)
        specs = build_stages(model, pp.split_points, pp.input_names)
        if len(specs) != pp.size:
            raise ValueError(f"model was cut into {len(specs)} stages but pp.size is {pp.size}")
        me = self.mesh.get_stage_id()
        self.spec = specs[me]
        log_rank0("pipeline: %d stages; stage %d recv=%s load=%s send=%s", len(specs), me, self.spec.recv_names,
                  self.spec.load_names, self.spec.send_names)
        # keep only this rank's stage; other stages' parameters are released
        keep = {id(p) for p in self.spec.module.parameters()}
        for i, s in enumerate(specs):
            if i != me:
                for p in s.module.parameters():
                    if id(p) not in keep:
                        p.data = torch.empty(0, dtype=p.dtype, device=p.device)
        self.model = self.spec.module
        self.executor = PipeExecutor(self.spec, self.model, self.mesh, self.device, pp.num_micro_batches,
                                     algo=pp.schedule, broadcast_loss=pp.broadcast_loss)

    def _get_underlay_model(self):
        return self.model

    def _update_underlay_model(self, model):
        self.model = model
        self.executor.module = model

    def _bind(self, args, kwargs):
        return bind_args_to_kwargs(args, kwargs, self.orig_forward_sig)

    def forward(self, *args, output_fn=None, **kwargs):
        """Evaluation: returns the outputs on the last stage, ``None`` elsewhere (reference pipeline.py:110-131)."""
        return self.executor.forward(self._bind(args, kwargs), output_fn)

    def forward_backward(self, *args, output_fn=None, **kwargs):
        """One training step's forward + backward over all micro-batches; returns the (averaged) loss
        (reference pipeline.py:133-149).  The optimizer step stays in the user loop."""
        return self.executor.forward_backward(self._bind(args, kwargs), output_fn)
