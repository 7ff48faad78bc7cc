"""Parallelism engine (reference torchacc/dist/).  Exposed both as ``torchacc_b200.parallel`` and, for API
compatibility, as ``torchacc_b200.dist``."""
from .bootstrap import (BACKEND_NAME, EAGER_BACKEND_NAME, backend_name, current_device, init_nccl_context,
                        init_process_group, local_rank, rank, rendezvous, world_size)
from .mesh import Mesh, ProcessTopology
from .parallel_module import ParallelModule
from .sharded import DataParallel, FullyShardedDataParallel, SpmdFullyShardedDataParallel
from .distributed_parallel import DistributedParallel
from . import fsdp, state_dict_utils, tp, pp
from .pp import PipelineParallel

__all__ = [
    "BACKEND_NAME", "EAGER_BACKEND_NAME", "backend_name", "current_device", "init_nccl_context", "init_process_group",
    "local_rank", "rank", "rendezvous", "world_size", "Mesh", "ProcessTopology", "ParallelModule", "DataParallel",
    "FullyShardedDataParallel", "SpmdFullyShardedDataParallel", "PipelineParallel", "DistributedParallel", "fsdp",
    "pp", "tp", "state_dict_utils",
]
