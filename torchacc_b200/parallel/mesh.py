"""Rank topology and process-group mesh.

Capability parity with reference torchacc/dist/mesh.py:13-418 (``ProcessTopology`` + ``Mesh`` and its
``get_{dp,pp,tp,fsdp}_{rank,num,proc_group,rank_groups}`` accessors), re-designed:

* the rank grid is a numpy ndarray indexed by axis name -- group enumeration is a moveaxis+reshape, not a
  hand-rolled cartesian walk;
* ``sp`` (sequence/context parallel) is a first-class axis (the reference stores ``sp_num`` but raises if
  'sp' appears in the topology, mesh.py:255-258), and the 2-D ulysses x ring split of the sp axis is derived
  here so context-parallel groups come from the same grid;
* process groups are created lazily per axis (``dist.new_group`` is a world-collective, so creation order is
  fixed: axis order of ``AXES``), and on NVSwitch boxes each group can carry a symmetric-memory domain.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch.distributed as dist

AXES = ("dp", "fsdp", "pp", "sp", "tp")


class ProcessTopology:
    """n-D cartesian map between global ranks and per-axis coordinates.

    ``axes`` lists axis names from slowest- to fastest-varying; ``dims`` are their sizes.
    """

    def __init__(self, axes: Sequence[str], dims: Sequence[int]):
        if len(axes) != len(dims):
            raise ValueError("axes and dims must have the same length")
        if len(set(axes)) != len(axes):
            raise ValueError(f"duplicate axis in {axes}")
        self.axes = list(axes)
        self.dims = [int(d) for d in dims]
        self.grid = np.arange(int(np.prod(self.dims)), dtype=np.int64).reshape(self.dims)

    def world_size(self) -> int:
        return int(self.grid.size)

    def get_dim(self, axis: str) -> int:
        return self.dims[self.axes.index(axis)] if axis in self.axes else 1

    def get_rank(self, **coords) -> int:
        idx = tuple(int(coords[a]) for a in self.axes)
        return int(self.grid[idx])

    def get_coord(self, rank: int) -> Dict[str, int]:
        idx = np.unravel_index(int(rank), self.dims)
        return {a: int(i) for a, i in zip(self.axes, idx)}

    def get_axis_comm_lists(self, axis: str) -> List[List[int]]:
        """All rank lists that differ only along ``axis`` (one communicator per list)."""
        if axis not in self.axes:
            return [[r] for r in range(self.world_size())]
        k = self.axes.index(axis)
        moved = np.moveaxis(self.grid, k, -1).reshape(-1, self.dims[k])
        return [list(map(int, row)) for row in moved]

    def get_multi_axis_comm_lists(self, axes: Sequence[str]) -> List[List[int]]:
        """Rank lists spanning several axes jointly (e.g. ('dp','fsdp') for the full data-parallel group).
        Ranks inside a list are ordered with the first axis slowest."""
        axes = [a for a in axes if a in self.axes]
        if not axes:
            return [[r] for r in range(self.world_size())]
        ks = [self.axes.index(a) for a in axes]
        moved = np.moveaxis(self.grid, ks, list(range(-len(ks), 0)))
        width = int(np.prod([self.dims[k] for k in ks]))
        return [list(map(int, row)) for row in moved.reshape(-1, width)]

    def filter_match(self, **coords) -> List[int]:
        """Ranks whose coordinates match all given ``axis=value`` pairs."""
        sl = tuple(int(coords[a]) if a in coords else slice(None) for a in self.axes)
        return sorted(map(int, np.asarray(self.grid[sl]).reshape(-1)))

    def get_axis_list(self, axis: str, idx: int) -> List[int]:
        return self.filter_match(**{axis: idx})

    def __repr__(self):
        return "ProcessTopology(" + ", ".join(f"{a}={d}" for a, d in zip(self.axes, self.dims)) + ")"


class Mesh:
    """Process-group mesh over dp / fsdp / pp / sp / tp.

    Args mirror the reference (mesh.py:230-239): ``dp_num, pp_num, tp_num, fsdp_num, sp_num, topology``.
    ``topology`` orders axes from slowest (inter-node) to fastest (adjacent ranks).
    """

    def __init__(self, dp_num: int = 1, pp_num: int = 1, tp_num: int = 1, fsdp_num: int = 1, sp_num: int = 1,
                 topology: Optional[List[str]] = None, sp_mode: str = "ulysses", ulysses_num: Optional[int] = None,
                 rank: Optional[int] = None, world_size: Optional[int] = None, create_groups: bool = True):
        sizes = {"dp": dp_num, "fsdp": fsdp_num, "pp": pp_num, "sp": sp_num, "tp": tp_num}
        for a, n in sizes.items():
            if not isinstance(n, int) or n < 1:
                raise ValueError(f"{a}_num must be a positive int, got {n!r}")
        topo = list(topology) if topology is not None else ["dp", "fsdp", "pp", "sp", "tp"]
        if len(set(topo)) != len(topo):
            raise ValueError("duplicate axis in topology")
        for t in topo:
            if t not in AXES:
                raise ValueError(f"unknown axis '{t}' in topology; expected a subset of {AXES}")
        for a in AXES:  # axes the caller omitted: size must be 1, or we slot them in before 'tp'
            if a not in topo:
                if sizes[a] != 1:
                    topo.insert(topo.index("tp") if "tp" in topo else len(topo), a)
                else:
                    topo.append(a)
        self.sizes = sizes
        self.topology = ProcessTopology(topo, [sizes[a] for a in topo])
        initialized = dist.is_available() and dist.is_initialized()
        self.world_size = world_size if world_size is not None else (dist.get_world_size() if initialized else 1)
        self.global_rank = rank if rank is not None else (dist.get_rank() if initialized else 0)
        if self.topology.world_size() != self.world_size:
            raise ValueError(f"parallel degrees {sizes} use {self.topology.world_size()} ranks but the world has "
                             f"{self.world_size}: the configured strategy must cover every device")
        self.coord = self.topology.get_coord(self.global_rank)
        self.sp_mode = sp_mode
        # 2-D context parallel: fast sub-axis = ulysses, slow sub-axis = ring
        if sp_mode == "ulysses":
            self.ulysses_num, self.ring_num = sp_num, 1
        elif sp_mode == "ring":
            self.ulysses_num, self.ring_num = 1, sp_num
        else:
            self.ulysses_num = ulysses_num if ulysses_num else sp_num
            if sp_num % self.ulysses_num:
                raise ValueError("ulysses_num must divide sp_num")
            self.ring_num = sp_num // self.ulysses_num
        self._rank_groups: Dict[str, List[List[int]]] = {}
        self._groups: Dict[str, object] = {}
        for a in AXES:
            self._rank_groups[a] = self.topology.get_axis_comm_lists(a)
        # joint data-parallel group (dp x fsdp): loss averaging, HSDP replicas, grad-norm
        self._rank_groups["data"] = self.topology.get_multi_axis_comm_lists(("dp", "fsdp"))
        # ranks holding the same parameter shard and seeing different tokens: dp x sp (gradient all-reduce group of
        # DP / HSDP; context-parallel ranks are data-parallel as far as parameters are concerned)
        self._rank_groups["replica"] = self.topology.get_multi_axis_comm_lists(("dp", "sp"))
        # sub-groups of the sp axis
        ul, ring = [], []
        for line in self._rank_groups["sp"]:
            arr = np.asarray(line).reshape(self.ring_num, self.ulysses_num)
            ul += [list(map(int, r)) for r in arr]
            ring += [list(map(int, c)) for c in arr.T]
        self._rank_groups["ulysses"], self._rank_groups["ring"] = ul, ring
        if create_groups and initialized and self.world_size > 1:
            self._create_groups()

    # ---- group construction ---------------------------------------------------------------------------
    def _create_groups(self):
        made = {}
        for name in ("dp", "fsdp", "pp", "sp", "tp", "data", "replica", "ulysses", "ring"):
            lists = self._rank_groups[name]
            if len(lists[0]) == 1:
                continue
            key = tuple(sorted(tuple(sorted(r)) for r in lists))
            if key in made:                      # identical partition already has communicators (e.g. replica == dp)
                if made[key] in self._groups:
                    self._groups[name] = self._groups[made[key]]
                continue
            made[key] = name
            if len(lists[0]) == self.world_size:
                self._groups[name] = dist.group.WORLD
                continue
            for ranks in lists:  # every rank must call new_group for every list, in the same order
                g = dist.new_group(ranks=ranks)
                if self.global_rank in ranks:
                    self._groups[name] = g
        # second, independent communicator along the pp axis for gradients flowing backwards.  NCCL executes the p2p
        # operations between two ranks of ONE communicator in issue order; with early-posted receives a stage's
        # pending irecv(activation i+1) would sit in front of its isend(grad i) while the neighbour's
        # isend(activation i+1) sits behind its irecv(grad i) -- a cyclic wait (observed: 2-stage 1F1B hang on GPUs).
        lists = self._rank_groups["pp"]
        if len(lists[0]) > 1:
            for ranks in lists:
                g = dist.new_group(ranks=ranks)
                if self.global_rank in ranks:
                    self._groups["pp_bwd"] = g

    def _my_ranks(self, name: str) -> List[int]:
        for ranks in self._rank_groups[name]:
            if self.global_rank in ranks:
                return ranks
        raise RuntimeError(f"rank {self.global_rank} not in any {name} group")

    def get_proc_group(self, name: str):
        """Process group of this rank along ``name`` (None when the axis has size 1)."""
        return self._groups.get(name)

    def get_rank_groups(self, name: str) -> List[List[int]]:
        return self._rank_groups[name]

    def get_group_ranks(self, name: str) -> List[int]:
        return self._my_ranks(name)

    def get_axis_rank(self, name: str) -> int:
        return self._my_ranks(name).index(self.global_rank)

    # ---- reference-compatible accessors (mesh.py:328-418) ---------------------------------------------
    def get_global_rank(self): return self.global_rank
    def get_world_size(self): return self.world_size

    def get_dp_rank(self): return self.coord["dp"]
    def get_dp_num(self): return self.sizes["dp"]
    def get_dp_proc_group(self): return self.get_proc_group("dp")
    def get_dp_rank_groups(self): return self._rank_groups["dp"]

    def get_pp_rank(self): return self.coord["pp"]
    def get_pp_num(self): return self.sizes["pp"]
    def get_pp_proc_group(self): return self.get_proc_group("pp")
    def get_pp_bwd_proc_group(self): return self.get_proc_group("pp_bwd") or self.get_proc_group("pp")
    def get_pp_rank_groups(self): return self._rank_groups["pp"]
    def get_stage_id(self): return self.coord["pp"]
    def is_first_stage(self): return self.coord["pp"] == 0
    def is_last_stage(self): return self.coord["pp"] == self.sizes["pp"] - 1

    def stage_to_global(self, stage_id: int) -> int:
        """Global rank of pipeline stage ``stage_id`` on this rank's pipeline (mesh.py:362-365)."""
        c = dict(self.coord)
        c["pp"] = stage_id % self.sizes["pp"]
        return self.topology.get_rank(**c)

    def get_tp_rank(self): return self.coord["tp"]
    def get_tp_num(self): return self.sizes["tp"]
    def get_tp_proc_group(self): return self.get_proc_group("tp")
    def get_tp_rank_groups(self): return self._rank_groups["tp"]

    def get_fsdp_rank(self): return self.coord["fsdp"]
    def get_fsdp_num(self): return self.sizes["fsdp"]
    def get_fsdp_proc_group(self): return self.get_proc_group("fsdp")
    def get_fsdp_rank_groups(self): return self._rank_groups["fsdp"]

    def get_sp_rank(self): return self.coord["sp"]
    def get_sp_num(self): return self.sizes["sp"]
    def get_sp_proc_group(self): return self.get_proc_group("sp")
    def get_sp_rank_groups(self): return self._rank_groups["sp"]
    def get_ulysses_proc_group(self): return self.get_proc_group("ulysses")
    def get_ring_proc_group(self): return self.get_proc_group("ring")

    def get_replica_proc_group(self): return self.get_proc_group("replica")
    def get_replica_num(self): return self.sizes["dp"] * self.sizes["sp"]
    def get_data_proc_group(self): return self.get_proc_group("data")
    def get_data_num(self): return self.sizes["dp"] * self.sizes["fsdp"]
    def get_data_rank(self): return self._my_ranks("data").index(self.global_rank)

    def __repr__(self):
        return f"Mesh(rank={self.global_rank}/{self.world_size}, {self.topology!r}, coord={self.coord})"
