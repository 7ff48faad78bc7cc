"""Pipeline parallelism (reference torchacc/dist/pp/)."""
from . import microbatch, p2p, partition, schedule
from .executor import PipeExecutor
from .pipeline import PipelineParallel, preprocess_config

__all__ = ["PipelineParallel", "PipeExecutor", "preprocess_config", "schedule", "microbatch", "p2p", "partition"]
