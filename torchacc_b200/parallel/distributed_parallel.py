"""Composition of the parallel strategies (reference torchacc/dist/distributed_parallel.py:19-111).

Order, outermost first: PP -> (TP / CP rewrite of the stage's modules) -> FSDP -> DP.  DP as a separate wrapper is
only used when there is no FSDP (with FSDP the dp axis becomes the HYBRID replica group of the same engine).
"""
from __future__ import annotations

import torch.nn as nn

from .parallel_module import ParallelModule
from .sharded import DataParallel, FullyShardedDataParallel


class DistributedParallel(ParallelModule):

    def __init__(self, model: nn.Module, config, orig_forward_sig=None, **kwargs):
        super().__init__(model, config)
        self.orig_forward_sig = orig_forward_sig
        d = config.dist
        self.pp_wrapper = None
        m = model
        if self.has_tp or self.has_sp:
            from .tp import parallelize_model
            m = parallelize_model(m, config, self.mesh)
        if self.has_pp:
            from .pp.pipeline import PipelineParallel
            m = PipelineParallel(m, config, orig_forward_sig=orig_forward_sig)
            self.pp_wrapper = m
            inner = m._get_underlay_model()
        else:
            inner = m
        # every distributed configuration runs on the flat-parameter engine (uniform optimizer / clipping /
        # checkpoint surface); with dp == fsdp == 1 it degenerates to a local engine without collectives
        wants_engine = True
        if self.has_fsdp:
            inner = FullyShardedDataParallel(inner, config)
        elif wants_engine:
            inner = DataParallel(inner, config)
        if self.has_pp:
            m._update_underlay_model(inner)
        else:
            m = inner
        self.model = m

    def _inner_engine_module(self):
        m = self.model
        if self.pp_wrapper is not None:
            m = self.pp_wrapper._get_underlay_model()
        return m

    @property
    def engine(self):
        return getattr(self._inner_engine_module(), "engine", None)

    def forward(self, *args, **kwargs):
        return self.model(*args, **kwargs)

    def forward_backward(self, *args, output_fn=None, **kwargs):
        if self.pp_wrapper is None:
            raise NotImplementedError("forward_backward requires pipeline parallelism (pp.size > 1)")
        return self.pp_wrapper.forward_backward(*args, output_fn=output_fn, **kwargs)

    def parameters(self, recurse: bool = True):
        return self._inner_engine_module().parameters(recurse)

    def named_parameters(self, *a, **k):
        return self._inner_engine_module().named_parameters(*a, **k)

    def zero_grad(self, set_to_none: bool = True):
        m = self._inner_engine_module()
        return m.zero_grad(set_to_none)

    def clip_grad_norm_(self, max_norm, norm_type=2.0):
        return self._inner_engine_module().clip_grad_norm_(max_norm, norm_type)

    # optimizer-state facade (reference distributed_parallel.py:85-111)
    def sharded_optim_state_dict(self, optim):
        return self._inner_engine_module().sharded_optim_state_dict(optim)

    def full_optim_state_dict(self, optim, **kwargs):
        return self._inner_engine_module().full_optim_state_dict(optim, **kwargs)

    def optim_state_dict_to_load(self, optim_state_dict, **kwargs):
        return self._inner_engine_module().optim_state_dict_to_load(optim_state_dict, **kwargs)
