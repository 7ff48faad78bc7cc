"""Runtime glue (reference torchacc/core/__init__.py:12-62).

The reference's functions here drive the XLA lazy tensor runtime (``lazy_device``, ``sync`` -> ``mark_step``).
There is no lazy runtime in this framework, so they keep their names with eager meanings:
``lazy_device()`` is the rank's CUDA device, ``sync()`` is an (optional) stream synchronisation.
"""
from __future__ import annotations

import torch

from . import amp
from .async_loader import AsyncLoader
from .dynamic import mark_dynamic


def lazy_device() -> torch.device:
    """The device ``accelerate`` places the model on (reference core/__init__.py:17-25)."""
    from ..parallel.bootstrap import current_device
    return current_device()


def is_lazy_device(device) -> bool:
    """Always False: there is no lazy (XLA) device type."""
    return False


def is_lazy_tensor(tensor) -> bool:
    return False


def fetch_gradients(optimizer) -> list:
    """All gradients held by ``optimizer`` (reference core/__init__.py:38-46); understands flat shards whose
    gradient lives in ``_tb_grad``."""
    grads = []
    for group in optimizer.param_groups:
        for p in group["params"]:
            g = getattr(p, "_tb_grad", None)
            if g is None:
                g = p.grad
            if g is not None:
                grads.append(g)
    return grads


def sync(wait: bool = False) -> None:
    """Reference: cut + launch the lazy graph (``xm.mark_step``).  Eager: nothing to cut; with ``wait=True`` block
    until the device finished all queued work."""
    if wait and torch.cuda.is_available():
        torch.cuda.synchronize()


def mark_step() -> None:
    sync(False)


def save(obj, path, master_only: bool = True, global_master: bool = False) -> None:
    """``torch.save`` with tensors moved to CPU first (reference ``ta.save = xm.save``).  With ``master_only`` only
    rank 0 writes."""
    import os
    from ..utils.utils import apply_to_tensors
    if master_only and int(os.environ.get("RANK", "0")) != 0:
        return
    torch.save(apply_to_tensors(lambda t: t.detach().cpu(), obj), path)


def send_cpu_data_to_device(data, device, non_blocking: bool = True):
    from ..utils.utils import apply_to_tensors
    return apply_to_tensors(lambda t: t.to(device, non_blocking=non_blocking), data)


__all__ = ["AsyncLoader", "amp", "lazy_device", "is_lazy_device", "is_lazy_tensor", "fetch_gradients", "sync",
           "mark_step", "save", "send_cpu_data_to_device", "mark_dynamic"]
