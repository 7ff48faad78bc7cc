"""Utilities (reference torchacc/utils/)."""
from . import checkpoint, cpu_offload, decompose, import_utils, logger as _logger_mod, metrics, patch, trace, utils
from .logger import logger

from .metrics import ThroughputMeter, memory_stats

__all__ = ["checkpoint", "cpu_offload", "decompose", "import_utils", "logger", "metrics", "patch", "trace", "utils",
           "ThroughputMeter", "memory_stats"]
