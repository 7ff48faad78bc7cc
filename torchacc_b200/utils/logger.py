"""Framework logger.  Level comes from ``ACC_LOG_LEVEL`` like the reference (torchacc/utils/logger.py:4-15);
records carry the rank so multi-process logs are attributable."""
import logging
import os
import sys

_LEVEL = os.environ.get("ACC_LOG_LEVEL", "INFO").upper()


class _RankFilter(logging.Filter):
    def filter(self, record):
        record.rank = os.environ.get("RANK", "0")
        return True


def _build() -> logging.Logger:
    lg = logging.getLogger("torchacc_b200")
    if lg.handlers:
        return lg
    lg.setLevel(getattr(logging, _LEVEL, logging.INFO))
    h = logging.StreamHandler(sys.stderr)
    h.setFormatter(logging.Formatter("[%(asctime)s r%(rank)s %(levelname)s] %(message)s", "%H:%M:%S"))
    h.addFilter(_RankFilter())
    lg.addHandler(h)
    lg.propagate = False
    return lg


logger = _build()


def log_rank0(msg, *args, level=logging.INFO):
    if os.environ.get("RANK", "0") == "0":
        logger.log(level, msg, *args)
