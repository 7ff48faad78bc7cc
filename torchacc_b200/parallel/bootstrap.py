"""Process bootstrap: env ranks, process-group init (NCCL on GPUs, gloo on CPU), rendezvous.

Reference: torchacc/dist/__init__.py:33-116 (``world_size/rank/local_rank/init_process_group/
init_nccl_context/rendezvous``).  The reference's eager path hard-wires 'nccl' (dist/backend.py:18) and cannot
run the CPU plumbing configuration; here the backend follows the device so every multi-process code path is
testable with gloo.
"""
from __future__ import annotations

import datetime
import os

import torch
import torch.distributed as dist

EAGER_BACKEND_NAME = "nccl"
BACKEND_NAME = "nccl"  # kept for API compatibility; resolved per device in ``backend_name()``


def world_size() -> int:
    return int(os.environ.get("WORLD_SIZE", "1"))


def rank() -> int:
    return int(os.environ.get("RANK", "0"))


def local_rank() -> int:
    return int(os.environ.get("LOCAL_RANK", "0"))


def use_cuda() -> bool:
    return torch.cuda.is_available() and os.environ.get("TORCHACC_B200_FORCE_CPU", "0") != "1"


def backend_name() -> str:
    return "nccl" if use_cuda() else "gloo"


def current_device() -> torch.device:
    if use_cuda():
        return torch.device("cuda", local_rank() % max(torch.cuda.device_count(), 1))
    return torch.device("cpu")


def init_process_group(config=None, timeout_s: int = 1800) -> None:
    """Idempotent ``torch.distributed`` initialisation from the torchrun environment."""
    from ..utils import watchdog
    watchdog.arm_from_env()            # TORCHACC_B200_HANG_DUMP=<s>: dump every thread's Python stack on a hang
    if not dist.is_available() or dist.is_initialized():
        return
    if world_size() == 1 and "MASTER_ADDR" not in os.environ:
        return  # single process: no process group needed
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    kwargs = dict(backend=backend_name(), rank=rank(), world_size=world_size(),
                  timeout=datetime.timedelta(seconds=timeout_s))
    if use_cuda():
        dev = current_device()
        torch.cuda.set_device(dev)
        try:
            dist.init_process_group(device_id=dev, **kwargs)
        except TypeError:  # older signature
            dist.init_process_group(**kwargs)
    else:
        dist.init_process_group(**kwargs)


def init_nccl_context(config=None) -> None:
    """Warm the communicators that the first training step would otherwise create lazily (the reference does
    this for PP p2p, dist/__init__.py:58-98).  One tiny all-reduce per existing group."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    dev = current_device()
    t = torch.ones(1, device=dev)
    dist.all_reduce(t)
    if config is not None:
        mesh = config.get_mesh()
        for name in ("dp", "fsdp", "pp", "sp", "tp"):
            g = mesh.get_proc_group(name)
            if g is not None and g is not dist.group.WORLD:
                dist.all_reduce(t, group=g)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)


def rendezvous(tag: str = "", payload: bytes = b"", replicas=None) -> None:
    """Barrier with a tag (reference dist/__init__.py:101-116 wraps xm.rendezvous)."""
    if dist.is_available() and dist.is_initialized():
        if use_cuda():
            dist.barrier(device_ids=[current_device().index])
        else:
            dist.barrier()
