"""Collective back-ends used by the sharding engines.

One small interface -- ``all_gather(shard, full)``, ``reduce_scatter(full, shard_out)``, ``all_reduce(t)`` -- with
three implementations:

* ``SymmCollectives``  : our own kernels over NVLink peer memory (csrc/comm): every rank's staging buffers are
  IPC-mapped into every peer, all-gather = peers' shards pulled with 16-byte loads, reduce-scatter = each rank
  reads its slice from all peers and accumulates in fp32 (bf16 on the wire), device-side barriers on signal pads;
* ``NcclCollectives``  : torch.distributed / NCCL (multi-node fallback and the baseline the fused path is measured
  against);
* ``GlooCollectives``  : CPU plumbing tier (gloo lacks reduce_scatter: all_reduce + slice).

The reference calls NCCL through torch FSDP / XLA for all of these.
"""
from __future__ import annotations


import torch
import torch.distributed as dist


class Collectives:
    name = "base"

    def __init__(self, group, world: int, rank: int):
        self.group, self.world, self.rank = group, world, rank

    def all_gather(self, shard: torch.Tensor, full: torch.Tensor) -> None:
        raise NotImplementedError

    def reduce_scatter(self, full: torch.Tensor, out: torch.Tensor, scale: float = 1.0) -> None:
        """out[fp32 or same dtype] = scale * sum_r full_r[rank*n:(rank+1)*n]"""
        raise NotImplementedError

    def all_reduce(self, t: torch.Tensor, scale: float = 1.0) -> None:
        raise NotImplementedError

    def all_to_all(self, inp: torch.Tensor, out: torch.Tensor) -> None:
        """Equal-chunk all-to-all on flat contiguous tensors."""
        dist.all_to_all_single(out, inp, group=self.group)

    def alloc(self, numel: int, dtype: torch.dtype, device=None) -> torch.Tensor:
        """Buffer that will be the *source* of a collective (symmetric memory for the peer-memory back-end)."""
        return torch.zeros(numel, dtype=dtype, device=device)


class LocalCollectives(Collectives):
    """world == 1."""
    name = "local"

    def all_gather(self, shard, full):
        if full.data_ptr() != shard.data_ptr():
            full.copy_(shard)

    def reduce_scatter(self, full, out, scale=1.0):
        if out.data_ptr() != full.data_ptr():
            out.copy_(full)
        if scale != 1.0:
            out.mul_(scale)

    def all_reduce(self, t, scale=1.0):
        if scale != 1.0:
            t.mul_(scale)

    def all_to_all(self, inp, out):
        out.copy_(inp)


class NcclCollectives(Collectives):
    name = "nccl"

    def all_gather(self, shard, full):
        dist.all_gather_into_tensor(full, shard, group=self.group)

    def reduce_scatter(self, full, out, scale=1.0):
        if out.dtype == full.dtype:
            dist.reduce_scatter_tensor(out, full, group=self.group)
            if scale != 1.0:
                out.mul_(scale)
        else:  # fp32 accumulation requested: upcast before the wire (what torch FSDP reduce_dtype=fp32 does)
            tmp = full.to(out.dtype)
            if scale != 1.0:
                tmp.mul_(scale)
            dist.reduce_scatter_tensor(out, tmp, group=self.group)

    def all_reduce(self, t, scale=1.0):
        if scale != 1.0:
            t.mul_(scale)
        dist.all_reduce(t, group=self.group)


class GlooCollectives(Collectives):
    name = "gloo"

    def all_to_all(self, inp, out):
        ins = list(inp.reshape(self.world, -1).unbind(0))
        outs = [torch.empty_like(c) for c in ins]
        if inp.dtype == torch.bfloat16:
            ins = [c.view(torch.float16).contiguous() for c in ins]
            outs = [c.view(torch.float16) for c in outs]
        work = []
        # gloo has no all_to_all: pairwise exchange
        for r in range(self.world):
            if r == self.rank:
                outs[r].copy_(ins[r])
        for shift in range(1, self.world):
            dst, src = (self.rank + shift) % self.world, (self.rank - shift) % self.world
            g_dst = dist.get_global_rank(self.group, dst) if self.group is not None else dst
            g_src = dist.get_global_rank(self.group, src) if self.group is not None else src
            sw = dist.isend(ins[dst].contiguous(), g_dst, group=self.group)
            dist.recv(outs[src], g_src, group=self.group)
            sw.wait()
        flat = torch.cat([o.reshape(-1) for o in outs])
        out.reshape(-1).view(flat.dtype).copy_(flat)

    def all_gather(self, shard, full):
        if full.dtype == torch.bfloat16:  # gloo has no bf16 kernels on every build: move as fp16 bit patterns (pure copies)
            # gathered straight into views of the destination (no torch.cat: it trips CPU autocast with fp16 inputs)
            parts = list(full.view(torch.float16).reshape(-1).chunk(self.world))
            dist.all_gather(parts, shard.view(torch.float16).contiguous().reshape(-1), group=self.group)
        else:
            parts = [torch.empty_like(shard) for _ in range(self.world)]
            dist.all_gather(parts, shard.contiguous(), group=self.group)
            full.copy_(torch.cat(parts))

    def reduce_scatter(self, full, out, scale=1.0):
        tmp = full.float()
        if scale != 1.0:
            tmp.mul_(scale)
        dist.all_reduce(tmp, group=self.group)
        n = out.numel()
        out.copy_(tmp[self.rank * n:(self.rank + 1) * n].to(out.dtype))

    def all_reduce(self, t, scale=1.0):
        if t.dtype == torch.bfloat16:
            tmp = t.float()
            if scale != 1.0:
                tmp.mul_(scale)
            dist.all_reduce(tmp, group=self.group)
            t.copy_(tmp.to(t.dtype))
        else:
            if scale != 1.0:
                t.mul_(scale)
            dist.all_reduce(t, group=self.group)


def make_collectives(group, device: torch.device, prefer_symm: bool = True) -> Collectives:
    """Pick the implementation for ``group`` (None => single process)."""
    if group is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return LocalCollectives(group, 1, 0)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if device.type != "cuda":
        return GlooCollectives(group, world, rank)
    if prefer_symm:
        try:
            from .symm_mem import SymmCollectives, symm_available
            if symm_available(group):
                return SymmCollectives(group, world, rank, device)
        except Exception as e:  # pragma: no cover - depends on the box
            from ..utils.logger import logger
            logger.warning("symmetric-memory collectives unavailable (%s); using NCCL", e)
    return NcclCollectives(group, world, rank)
