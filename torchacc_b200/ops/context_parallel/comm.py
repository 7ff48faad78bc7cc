"""Differentiable collectives for context parallelism (reference torchacc/ops/context_parallel/utils.py:140-299).

* ``split_forward_gather_backward`` / ``gather_forward_split_backward`` along the sequence dim -- ``grad_scale`` is
  honoured (the reference accepts it and drops it);
* ``seq_head_all_to_all``: the Ulysses exchange, scatter one dim / gather another, implemented as ONE equal-chunk
  all-to-all on a packed buffer (the reference chunks into a python list, ``dist.all_to_all`` + ``cat``).
  On an NVSwitch domain the all-to-all is our peer-memory kernel (parallel/symm_mem.py)."""
from __future__ import annotations


import torch
import torch.distributed as dist

_COLL = {}


def _coll(group, device):
    from ...parallel.collectives import make_collectives
    key = (id(group), device.type)
    if key not in _COLL:
        _COLL[key] = make_collectives(group, device)
    return _COLL[key]


def _world(group) -> int:
    return dist.get_world_size(group) if (group is not None and dist.is_initialized()) else 1


def _rank(group) -> int:
    return dist.get_rank(group) if (group is not None and dist.is_initialized()) else 0


def _all_gather_dim(x: torch.Tensor, dim: int, group) -> torch.Tensor:
    w = _world(group)
    if w == 1:
        return x
    xm = x.movedim(dim, 0).contiguous()
    full = torch.empty((w * xm.shape[0], *xm.shape[1:]), dtype=x.dtype, device=x.device)
    _coll(group, x.device).all_gather(xm.reshape(-1), full.reshape(-1))
    return full.movedim(0, dim)


def _split_dim(x: torch.Tensor, dim: int, group) -> torch.Tensor:
    w, r = _world(group), _rank(group)
    if w == 1:
        return x
    assert x.shape[dim] % w == 0, f"dim {dim} of size {x.shape[dim]} is not divisible by the group size {w}"
    return x.chunk(w, dim=dim)[r].contiguous()


def _scale(g, grad_scale, group):
    if grad_scale == "up":
        return g * _world(group)
    if grad_scale == "down":
        return g / _world(group)
    return g


class _SplitFwdGatherBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, group, grad_scale):
        ctx.dim, ctx.group, ctx.grad_scale = dim, group, grad_scale
        return _split_dim(x, dim, group)

    @staticmethod
    def backward(ctx, g):
        return _all_gather_dim(_scale(g, ctx.grad_scale, ctx.group), ctx.dim, ctx.group), None, None, None


class _GatherFwdSplitBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, group, grad_scale):
        ctx.dim, ctx.group, ctx.grad_scale = dim, group, grad_scale
        return _all_gather_dim(x, dim, group)

    @staticmethod
    def backward(ctx, g):
        return _split_dim(_scale(g, ctx.grad_scale, ctx.group), ctx.dim, ctx.group), None, None, None


def split_forward_gather_backward(tensor, seq_dim, process_group, grad_scale=None):
    return _SplitFwdGatherBwd.apply(tensor, seq_dim, process_group, grad_scale)


def gather_forward_split_backward(tensor, seq_dim, process_group, grad_scale=None):
    return _GatherFwdSplitBwd.apply(tensor, seq_dim, process_group, grad_scale)


def _a2a(x: torch.Tensor, scatter_dim: int, gather_dim: int, group) -> torch.Tensor:
    """Split ``scatter_dim`` over the group, concatenate the received pieces along ``gather_dim``."""
    w = _world(group)
    if w == 1:
        return x
    assert x.shape[scatter_dim] % w == 0
    # [w, ..., scatter/w, ...] with the destination rank as the leading dim
    parts = torch.stack(x.chunk(w, dim=scatter_dim), 0).contiguous()
    out = torch.empty_like(parts)
    _coll(group, x.device).all_to_all(parts.reshape(-1), out.reshape(-1))
    return torch.cat(list(out.unbind(0)), dim=gather_dim)


class _AllToAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scatter_dim, gather_dim, group):
        ctx.cfg = (scatter_dim, gather_dim, group)
        return _a2a(x, scatter_dim, gather_dim, group)

    @staticmethod
    def backward(ctx, g):
        s, gd, group = ctx.cfg
        return _a2a(g.contiguous(), gd, s, group), None, None, None


def seq_head_all_to_all(x, scatter_dim: int, gather_dim: int, group):
    """Differentiable all-to-all (reference ``diff_all_to_all``, utils.py:275-299)."""
    return _AllToAll.apply(x, scatter_dim, gather_dim, group)


diff_all_to_all = seq_head_all_to_all
all_gather = _all_gather_dim
