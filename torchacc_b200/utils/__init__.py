"""Utilities (reference torchacc/utils/)."""
from . import checkpoint, cpu_offload, decompose, import_utils, logger as _logger_mod, patch, trace, utils
from .logger import logger

__all__ = ["checkpoint", "cpu_offload", "decompose", "import_utils", "logger", "patch", "trace", "utils"]
