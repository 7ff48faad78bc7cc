"""Deadlock analysis of the pipeline schedules under NCCL's point-to-point semantics.

gloo completes isend/irecv independently, NCCL does not: the p2p operations two ranks exchange over ONE communicator are
executed in issue order on each side, so the heads of both queues must be a matching send/recv pair.  With receives
posted early (our schedules do that to overlap transfers with compute) a single communicator deadlocks the 2-stage
1F1B schedule -- observed on 2 GPUs -- which is why activations and gradients travel on separate process groups
(parallel/pp/p2p.py).  This test replays the real instruction streams of every stage against that queue model."""
import itertools

import pytest

from torchacc_b200.parallel.pp import schedule as S


def _simulate(stages, micro, algo, split_channels):
    """Returns True if every rank finishes, False on deadlock."""
    progs = []
    for s in range(stages):
        sch = S.create_scheduler(algo, True, micro, stages, s)
        progs.append([i for step in sch.steps() for i in step])
    pc = [0] * stages
    queues = {}          # (rank, peer, channel) -> list of [kind, tag, done]
    waiting = [None] * stages     # op a rank is blocked on
    first_send = [True] * stages
    first_recv = [True] * stages
    pending_sends = [[] for _ in range(stages)]

    def chan(kind):
        return ("fwd" if kind == "act" else "bwd") if split_channels else "one"

    def enqueue(rank, peer, kind, direction, tag):
        op = [direction, tag, False]
        queues.setdefault((rank, peer, chan(kind)), []).append(op)
        return op

    def match():
        progressed = False
        for (a, b, c), qa in list(queues.items()):
            if a > b:
                continue
            qb = queues.get((b, a, c), [])
            while qa and qb:
                x, y = qa[0], qb[0]
                if {x[0], y[0]} == {"send", "recv"}:
                    assert x[1] == y[1], ("message order mismatch", x, y)
                    x[2] = y[2] = True
                    qa.pop(0); qb.pop(0)
                    progressed = True
                else:
                    break           # send/send or recv/recv at the heads: NCCL stalls this communicator
        return progressed

    recv_ops = [dict() for _ in range(stages)]
    while True:
        progressed = False
        for r in range(stages):
            while pc[r] < len(progs[r]):
                if waiting[r] is not None:
                    if not waiting[r][2]:
                        break
                    waiting[r] = None
                    progressed = True
                    continue
                ins = progs[r][pc[r]]
                name = type(ins).__name__
                if name == "SendActivation":
                    if first_send[r]:                      # blocking header send before the first tensors
                        first_send[r] = False
                        waiting[r] = enqueue(r, r + 1, "act", "send", ("hdr",))
                        continue                            # re-visit this instruction after the header completed
                    pending_sends[r].append(enqueue(r, r + 1, "act", "send", ("act", ins.micro_batch)))
                elif name == "PostRecvActivation":
                    if first_recv[r]:
                        first_recv[r] = False
                        waiting[r] = enqueue(r, r - 1, "act", "recv", ("hdr",))
                        continue
                    recv_ops[r][("act", ins.micro_batch)] = enqueue(r, r - 1, "act", "recv", ("act", ins.micro_batch))
                elif name == "WaitRecvActivation":
                    op = recv_ops[r][("act", ins.micro_batch)]
                    if not op[2]:
                        waiting[r] = op
                        pc[r] += 1
                        progressed = True
                        continue
                elif name == "SendGrad":
                    pending_sends[r].append(enqueue(r, r - 1, "grad", "send", ("grad", ins.micro_batch)))
                elif name == "PostRecvGrad":
                    recv_ops[r][("grad", ins.micro_batch)] = enqueue(r, r + 1, "grad", "recv", ("grad", ins.micro_batch))
                elif name == "WaitRecvGrad":
                    op = recv_ops[r][("grad", ins.micro_batch)]
                    if not op[2]:
                        waiting[r] = op
                        pc[r] += 1
                        progressed = True
                        continue
                pc[r] += 1
                progressed = True
        if match():
            progressed = True
        done = all(pc[r] >= len(progs[r]) and waiting[r] is None for r in range(stages)) and \
            all(op[2] for r in range(stages) for op in pending_sends[r])
        if done:
            return True
        if not progressed:
            return False


@pytest.mark.parametrize("stages,micro", list(itertools.product((2, 3, 4), (1, 2, 4, 8))))
@pytest.mark.parametrize("algo", ["1f1b", "gpipe"])
def test_schedules_do_not_deadlock_with_split_channels(stages, micro, algo):
    assert _simulate(stages, micro, algo, split_channels=True)


def test_single_communicator_would_deadlock_1f1b():
    """The configuration that hung on hardware: 2 stages, early-posted receives, one communicator."""
    assert not _simulate(2, 4, "1f1b", split_channels=False)
