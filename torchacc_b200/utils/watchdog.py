"""Hang diagnostics (the reference has none -- a lost peer is a silent hang until the c10d timeout).

``arm(seconds)`` makes the process dump the Python stack of every thread to stderr when it is still inside the armed
region after ``seconds`` (``faulthandler``: works while the main thread is blocked inside a CUDA / NCCL call), and
``disarm()`` cancels it.  ``TORCHACC_B200_HANG_DUMP=<seconds>`` arms it for the whole process from
``init_process_group`` (repeating), which is how the multi-GPU tests and examples are run on hardware.
``StepWatchdog`` is the per-step form used by training loops: ``with StepWatchdog(120): step()``.
"""
from __future__ import annotations

import faulthandler
import os
import sys

_armed = False


def arm(seconds: float, repeat: bool = False, exit: bool = False) -> None:
    global _armed
    faulthandler.dump_traceback_later(float(seconds), repeat=repeat, file=sys.stderr, exit=exit)
    _armed = True


def disarm() -> None:
    global _armed
    if _armed:
        faulthandler.cancel_dump_traceback_later()
        _armed = False


def arm_from_env() -> None:
    v = os.environ.get("TORCHACC_B200_HANG_DUMP", "")
    if v:
        try:
            arm(float(v), repeat=True)
        except ValueError:
            pass


class StepWatchdog:
    """Context manager: dump all Python stacks if the body runs longer than ``seconds`` (and optionally abort)."""

    def __init__(self, seconds: float, exit: bool = False):
        self.seconds, self.exit = seconds, exit

    def __enter__(self):
        arm(self.seconds, repeat=False, exit=self.exit)
        return self

    def __exit__(self, *exc):
        disarm()
        return False
