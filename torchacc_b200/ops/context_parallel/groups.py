"""Context-parallel process groups (reference torchacc/ops/context_parallel/init_group.py:20-112).

``initialize_context_parallel(cp_size, intra_size)`` lays the world out as ``[dp, inter, intra]`` (intra fastest):
Ulysses all-to-all runs inside an ``intra`` group, ring attention across the ``inter`` group, and the flat
``cp`` group spans both.  When the framework's ``Mesh`` already exists its ``sp / ulysses / ring`` groups are reused
(the reference builds a second, independent set of groups here)."""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch.distributed as dist

_GROUPS: Dict[str, object] = {}
_SIZES: Dict[str, int] = {}


def initialize_parallel_group(sizes: List[int]) -> List[object]:
    """n-D reshape of the world ranks; returns this rank's group along every axis (slowest axis first)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if int(np.prod(sizes)) != world:
        raise ValueError(f"group sizes {sizes} do not multiply to the world size {world}")
    grid = np.arange(world).reshape(sizes)
    mine = []
    for ax in range(len(sizes)):
        lines = np.moveaxis(grid, ax, -1).reshape(-1, sizes[ax])
        g_mine = None
        for line in lines:
            ranks = [int(x) for x in line]
            g = dist.new_group(ranks) if len(ranks) > 1 else None
            if rank in ranks:
                g_mine = g
        mine.append(g_mine)
    return mine


def initialize_context_parallel(context_parallel_size: int, intra_parallel_size: Optional[int] = None) -> None:
    world = dist.get_world_size()
    cp = context_parallel_size
    intra = intra_parallel_size if intra_parallel_size is not None else cp
    if world % cp or cp % intra:
        raise ValueError("context_parallel_size must divide the world size and be a multiple of intra_parallel_size")
    inter = cp // intra
    dp_g, inter_g, intra_g = initialize_parallel_group([world // cp, inter, intra])
    # flat cp group = (inter, intra) jointly
    rank = dist.get_rank()
    grid = np.arange(world).reshape(world // cp, cp)
    cp_g = None
    for line in grid:
        ranks = [int(x) for x in line]
        g = dist.new_group(ranks) if len(ranks) > 1 else None
        if rank in ranks:
            cp_g = g
    _GROUPS.update(dp=dp_g, inter=inter_g, intra=intra_g, cp=cp_g)
    _SIZES.update(cp=cp, inter=inter, intra=intra)


def use_mesh(mesh) -> None:
    """Adopt the sp / ulysses / ring groups of a framework ``Mesh``."""
    _GROUPS.update(cp=mesh.get_sp_proc_group(), intra=mesh.get_ulysses_proc_group(), inter=mesh.get_ring_proc_group(),
                   dp=mesh.get_dp_proc_group())
    _SIZES.update(cp=mesh.get_sp_num(), intra=mesh.ulysses_num, inter=mesh.ring_num)


def get_context_parallel_group():
    return _GROUPS.get("cp")


def get_intra_cp_process_group():
    return _GROUPS.get("intra")


def get_inter_cp_process_group():
    return _GROUPS.get("inter")


def get_dp_process_group():
    return _GROUPS.get("dp")


def get_context_parallel_size() -> int:
    return _SIZES.get("cp", 1)
