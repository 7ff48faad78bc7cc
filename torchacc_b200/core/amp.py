"""Loss scaling for fp16 training (reference torchacc/core/amp.py:9-42).

The reference subclasses torch_xla's syncfree GradScaler and all-reduces ``found_inf`` over the pipeline group.
This one works on the sharding engine's flat gradient shards: unscale + non-finite check is one pass of our
``sqnorm`` kernel per shard (device-resident flags, no host sync), ``found_inf`` is max-reduced over every group
that holds a different part of the model (fsdp shards, tp, pp), and ``FusedAdamW`` skips the update on the device.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


class GradScaler:

    def __init__(self, init_scale: float = 2.0 ** 16, growth_factor: float = 2.0, backoff_factor: float = 0.5,
                 growth_interval: int = 2000, enabled: bool = True, use_zero_grad: bool = False):
        self._enabled = enabled
        self._init_scale = init_scale
        self._growth_factor, self._backoff_factor = growth_factor, backoff_factor
        self._growth_interval = growth_interval
        self._scale: Optional[torch.Tensor] = None
        self._growth_tracker: Optional[torch.Tensor] = None
        self._found_inf: Optional[torch.Tensor] = None
        self._unscaled = set()
        self.use_zero_grad = use_zero_grad

    # ---- helpers ----------------------------------------------------------------------------------------
    def _lazy_init(self, device):
        if self._scale is None:
            self._scale = torch.full((1,), self._init_scale, dtype=torch.float32, device=device)
            self._growth_tracker = torch.zeros(1, dtype=torch.int32, device=device)

    def is_enabled(self):
        return self._enabled

    def get_scale(self) -> float:
        return float(self._scale) if self._scale is not None else self._init_scale

    def scale(self, outputs):
        if not self._enabled:
            return outputs
        from ..utils.utils import apply_to_tensors
        first = outputs if isinstance(outputs, torch.Tensor) else next(
            t for t in _iter_tensors(outputs))
        self._lazy_init(first.device)
        return apply_to_tensors(lambda t: t * self._scale.to(t.dtype), outputs)

    def _groups(self):
        from .. import get_global_context
        mesh = get_global_context().mesh
        if mesh is None or not dist.is_initialized():
            return []
        return [g for g in (mesh.get_fsdp_proc_group(), mesh.get_tp_proc_group(), mesh.get_pp_proc_group())
                if g is not None]

    def unscale_(self, optimizer):
        if not self._enabled or id(optimizer) in self._unscaled:
            return
        from ..ops.optim import grad_sqnorm, scale_
        from . import fetch_gradients
        grads = fetch_gradients(optimizer)
        device = grads[0].device if grads else self._scale.device
        self._lazy_init(device)
        inv = (1.0 / self._scale).float()
        for g in grads:
            scale_(g, inv)
        stat = grad_sqnorm(grads, device=device)
        found = (stat[1] > 0).float().reshape(1)
        found = torch.maximum(found, (~torch.isfinite(stat[0])).float().reshape(1))
        for grp in self._groups():   # reference amp.py:34-42 reduces over the pp group only
            dist.all_reduce(found, op=dist.ReduceOp.MAX, group=grp)
        self._found_inf = found
        self._unscaled.add(id(optimizer))

    def step(self, optimizer, *args, **kwargs):
        if not self._enabled:
            return optimizer.step(*args, **kwargs)
        self.unscale_(optimizer)
        if hasattr(optimizer, "found_inf") and hasattr(optimizer, "grad_scale"):
            optimizer.found_inf = self._found_inf      # device-side skip, no host sync
            out = optimizer.step(*args, **kwargs)
            optimizer.found_inf = None
            return out
        if float(self._found_inf) == 0.0:
            return optimizer.step(*args, **kwargs)
        return None

    def update(self, new_scale=None):
        if not self._enabled:
            return
        if new_scale is not None:
            self._scale.fill_(float(new_scale))
        elif self._found_inf is not None:
            found = self._found_inf
            grow = self._growth_tracker + 1
            self._scale = torch.where(found > 0, self._scale * self._backoff_factor,
                                      torch.where(grow >= self._growth_interval, self._scale * self._growth_factor,
                                                  self._scale))
            self._growth_tracker = torch.where(found > 0, torch.zeros_like(grow),
                                               torch.where(grow >= self._growth_interval, torch.zeros_like(grow), grow))
        self._found_inf = None
        self._unscaled.clear()

    def state_dict(self):
        return {"scale": self.get_scale(), "growth_factor": self._growth_factor,
                "backoff_factor": self._backoff_factor, "growth_interval": self._growth_interval,
                "_growth_tracker": int(self._growth_tracker) if self._growth_tracker is not None else 0}

    def load_state_dict(self, sd):
        self._init_scale = sd["scale"]
        self._growth_factor, self._backoff_factor = sd["growth_factor"], sd["backoff_factor"]
        self._growth_interval = sd["growth_interval"]
        if self._scale is not None:
            self._scale.fill_(sd["scale"])
            self._growth_tracker.fill_(sd["_growth_tracker"])


def _iter_tensors(obj):
    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, dict):
        for v in obj.values():
            yield from _iter_tensors(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _iter_tensors(v)
