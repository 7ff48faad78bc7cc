"""Pipeline schedules as instruction streams (reference torchacc/dist/pp/schedule.py:10-408).

A schedule is a generator of *steps*; each step is a list of instructions the executor runs in order.
Provided: ``OneFOneBTrain`` (1F1B / PipeDream-flush: warm-up forwards, steady 1F1B, cool-down backwards),
``GPipeTrain`` (all forwards then all backwards; not in the reference) and ``ForwardOnly`` (inference).
Unlike the reference the receive instructions are *posted early* (``PostRecv*``) and waited on right before use
(``WaitRecv*``) so NCCL p2p overlaps with the neighbouring micro-batch's compute.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List


@dataclass(frozen=True)
class Instr:
    micro_batch: int = -1
    buffer: int = -1

    @property
    def name(self):
        return type(self).__name__

    def __repr__(self):
        return f"{self.name}(mb={self.micro_batch}, buf={self.buffer})"


class LoadMicroBatch(Instr): pass
class ForwardPass(Instr): pass
class BackwardPass(Instr): pass
class SendActivation(Instr): pass
class RecvActivation(Instr): pass      # = post + wait (blocking form, kept for API parity)
class SendGrad(Instr): pass
class RecvGrad(Instr): pass
class PostRecvActivation(Instr): pass
class WaitRecvActivation(Instr): pass
class PostRecvGrad(Instr): pass
class WaitRecvGrad(Instr): pass
class ReduceGrads(Instr): pass
class ReduceTiedGrads(Instr): pass
class OptimizerStep(Instr): pass


class PipeSchedule:
    """Base: knows the stage position and the number of micro-batches."""

    def __init__(self, micro_batches: int, stages: int, stage_id: int):
        self.micro_batches, self.stages, self.stage_id = micro_batches, stages, stage_id
        self.prev_stage, self.next_stage = stage_id - 1, stage_id + 1

    @property
    def is_first_stage(self): return self.stage_id == 0
    @property
    def is_last_stage(self): return self.stage_id == self.stages - 1

    def num_pipe_buffers(self) -> int:
        return self.micro_batches

    def steps(self) -> Iterator[List[Instr]]:
        raise NotImplementedError

    def __iter__(self):
        return self.steps()

    def _buf(self, mb: int) -> int:
        return mb % self.num_pipe_buffers()


class ForwardOnly(PipeSchedule):
    """Inference: forward every micro-batch (reference ``PipeDreamFlushInfer``, schedule.py:122-153)."""

    def num_pipe_buffers(self):
        return 2

    def steps(self):
        for mb in range(self.micro_batches):
            b = self._buf(mb)
            cmds: List[Instr] = []
            if self.is_first_stage:
                cmds.append(LoadMicroBatch(mb, b))
            else:
                cmds += [PostRecvActivation(mb, b), WaitRecvActivation(mb, b)]
            if self.is_last_stage:
                cmds.append(LoadMicroBatch(mb, b))
            cmds.append(ForwardPass(mb, b))
            if not self.is_last_stage:
                cmds.append(SendActivation(mb, b))
            yield cmds


class OneFOneBTrain(PipeSchedule):
    """1F1B (reference ``PipeDreamFlushTrain``, schedule.py:156-227)."""

    def num_pipe_buffers(self):
        return max(min(self.stages - self.stage_id, self.micro_batches), 2)

    def steps(self):
        M, S, s = self.micro_batches, self.stages, self.stage_id
        warmup = min(S - s - 1, M)
        fwd = bwd = 0

        def forward_cmds(mb, post_next_grad):
            b = self._buf(mb)
            c: List[Instr] = []
            if not self.is_first_stage:
                c.append(WaitRecvActivation(mb, b))
            if self.is_first_stage or self.is_last_stage:
                c.append(LoadMicroBatch(mb, b))
            c.append(ForwardPass(mb, b))
            if not self.is_last_stage:
                c.append(SendActivation(mb, b))
            return c

        def backward_cmds(mb):
            b = self._buf(mb)
            c: List[Instr] = []
            if not self.is_last_stage:
                c.append(WaitRecvGrad(mb, b))
            c.append(BackwardPass(mb, b))
            if not self.is_first_stage:
                c.append(SendGrad(mb, b))
            return c

        # receives are posted one micro-batch ahead of their use
        if not self.is_first_stage and M > 0:
            yield [PostRecvActivation(0, self._buf(0))]
        for _ in range(warmup):
            c = forward_cmds(fwd, False)
            fwd += 1
            if not self.is_first_stage and fwd < M:
                c.insert(0, PostRecvActivation(fwd, self._buf(fwd)))
            yield c
        if not self.is_last_stage and M > 0:
            yield [PostRecvGrad(0, self._buf(0))]
        while fwd < M:                                   # steady state: one forward, one backward
            c = forward_cmds(fwd, True)
            fwd += 1
            if not self.is_first_stage and fwd < M:
                c.insert(0, PostRecvActivation(fwd, self._buf(fwd)))
            c += backward_cmds(bwd)
            bwd += 1
            if not self.is_last_stage and bwd < M:
                c.append(PostRecvGrad(bwd, self._buf(bwd)))
            yield c
        while bwd < M:                                   # cool-down
            c = backward_cmds(bwd)
            bwd += 1
            if not self.is_last_stage and bwd < M:
                c.append(PostRecvGrad(bwd, self._buf(bwd)))
            yield c
        yield [ReduceGrads(), OptimizerStep()]


class GPipeTrain(PipeSchedule):
    """All forwards, then all backwards (more activation memory, same bubble as 1F1B)."""

    def steps(self):
        M = self.micro_batches
        for mb in range(M):
            b = self._buf(mb)
            c: List[Instr] = []
            if not self.is_first_stage:
                c += [PostRecvActivation(mb, b), WaitRecvActivation(mb, b)]
            if self.is_first_stage or self.is_last_stage:
                c.append(LoadMicroBatch(mb, b))
            c.append(ForwardPass(mb, b))
            if not self.is_last_stage:
                c.append(SendActivation(mb, b))
            yield c
        for mb in range(M):
            b = self._buf(mb)
            c = []
            if not self.is_last_stage:
                c += [PostRecvGrad(mb, b), WaitRecvGrad(mb, b)]
            c.append(BackwardPass(mb, b))
            if not self.is_first_stage:
                c.append(SendGrad(mb, b))
            yield c
        yield [ReduceGrads(), OptimizerStep()]


# reference class names
PipeDreamFlushTrain = OneFOneBTrain
PipeDreamFlushInfer = ForwardOnly


class Algo:
    PipeDreamFlush = "1f1b"
    GPipe = "gpipe"


def create_scheduler(algo, training: bool, micro_batches: int, stages: int, stage_id: int) -> PipeSchedule:
    if not training:
        return ForwardOnly(micro_batches, stages, stage_id)
    if algo in (Algo.PipeDreamFlush, "1f1b", None):
        return OneFOneBTrain(micro_batches, stages, stage_id)
    if algo in (Algo.GPipe, "gpipe"):
        return GPipeTrain(micro_batches, stages, stage_id)
    raise ValueError(f"unknown pipeline schedule {algo!r}")


# instruction base-class names of the reference (torchacc/dist/pp/schedule.py:230,278)
PipeInstruction = Instr
BufferOpInstruction = Instr
