"""Pipeline executor: interprets a schedule's instruction stream for one stage
(reference torchacc/dist/pp/executor.py:16-725)."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

import torch
import torch.distributed as dist

from . import schedule as S
from .p2p import StageLink, flatten_values, unflatten_values
from .partition import StageSpec


def _first_loss(out):
    if isinstance(out, torch.Tensor):
        return out
    if isinstance(out, dict):
        if "loss" in out and out["loss"] is not None:
            return out["loss"]
    if hasattr(out, "loss") and out.loss is not None:
        return out.loss
    if isinstance(out, (tuple, list)) and out:
        return out[0]
    raise ValueError("the last pipeline stage must produce a loss (tensor, .loss, ['loss'] or first tuple element)")


class PipeExecutor:

    def __init__(self, spec: StageSpec, stage_module, mesh, device, num_micro_batches: int, algo: str = "1f1b",
                 broadcast_loss: bool = True):
        self.spec, self.module = spec, stage_module
        self.mesh, self.device = mesh, device
        self.micro_batches = num_micro_batches
        self.algo = algo
        self.broadcast_loss = broadcast_loss
        self.link = StageLink(mesh, device)
        self.stage, self.stages = mesh.get_stage_id(), mesh.get_pp_num()
        self.is_first, self.is_last = mesh.is_first_stage(), mesh.is_last_stage()
        # Receives are posted one micro-batch early so transfers overlap compute -- except during the FIRST schedule
        # run: CUDA loads kernel modules lazily, a first-use load can wait for running kernels, and an early-posted NCCL
        # receive is a kernel that runs until its peer sends.  The first run therefore posts each receive right before
        # its wait (the reference's blocking order, executor.py:622-667); every kernel is resident afterwards.
        import os as _os
        self.early_post = _os.environ.get("TORCHACC_B200_PP_EARLY_POST", "1") != "0"
        self._warm = False
        self._reset_state()

    def _reset_state(self):
        self.mb_kwargs: List[dict] = []
        self.recv_bufs: Dict[int, Any] = {}
        self.recv_works: Dict[int, list] = {}
        self.grad_bufs: Dict[int, Any] = {}
        self.grad_works: Dict[int, list] = {}
        self.inputs: Dict[int, List[torch.Tensor]] = {}
        self.outputs: Dict[int, List[torch.Tensor]] = {}
        self.loaded: Dict[int, dict] = {}
        self.results: Dict[int, Any] = {}
        self.total_loss: Optional[torch.Tensor] = None
        self.output_fn: Optional[Callable] = None
        self.output_fn_kwargs: List[dict] = []

    def reset_activation_shape(self):
        """Call when the micro-batch shapes change (reference executor.py:165-172)."""
        self.link.flush()
        self.link.reset_shapes()

    # ---- instruction handlers -----------------------------------------------------------------------------------
    def _load_micro_batch(self, ins):
        kw = self.mb_kwargs[ins.micro_batch]
        dev = self.device
        self.loaded[ins.buffer] = {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v)
                                   for k, v in kw.items() if k in self.spec.load_names}

    def _post_recv_act(self, ins):
        if self._warm and self.early_post:
            self.recv_bufs[ins.buffer], self.recv_works[ins.buffer] = self.link.post_recv_activations()

    def _wait_recv_act(self, ins):
        if ins.buffer not in self.recv_bufs:      # just-in-time posting (first run / early_post off)
            self.recv_bufs[ins.buffer], self.recv_works[ins.buffer] = self.link.post_recv_activations()
        StageLink.wait(self.recv_works.pop(ins.buffer, []))
        bufs = self.recv_bufs.pop(ins.buffer)
        for b, m in zip(bufs, self.link.recv_meta):
            if m.requires_grad and torch.is_grad_enabled():
                b.requires_grad_(True)
        self.inputs[ins.buffer] = bufs

    def _forward(self, ins):
        values = dict(self.loaded.pop(ins.buffer, {}))
        if not self.is_first:
            # boundary values are tensors or flat tuples of tensors (regrouped from the flat transfer list)
            for name, t in zip(self.spec.recv_names, unflatten_values(self.inputs[ins.buffer], self.link.recv_meta)):
                values[name] = t
        out = self.module(**values)
        if self.is_last:
            if self.spec.rebuild_output is not None and isinstance(out, dict):
                out = self.spec.rebuild_output(out)
            if self.output_fn is not None:
                out = self.output_fn(out, **self.output_fn_kwargs[ins.micro_batch])
            if torch.is_grad_enabled():
                loss = _first_loss(out)
                self.results[ins.buffer] = loss
                det = loss.detach().float()
                self.total_loss = det if self.total_loss is None else self.total_loss + det
            else:
                self.results[ins.micro_batch] = out
        else:
            if not isinstance(out, dict):
                if len(self.spec.send_names) != 1:
                    raise ValueError("a stage with several outgoing values must return a dict")
                out = {self.spec.send_names[0]: out}
            self.outputs[ins.buffer], self._send_groups = flatten_values([out[n] for n in self.spec.send_names])

    def _send_act(self, ins):
        self.link.send_activations(self.outputs[ins.buffer], self._send_groups)
        if not torch.is_grad_enabled():
            self.outputs.pop(ins.buffer, None)

    def _post_recv_grad(self, ins):
        if self._warm and self.early_post:
            self.grad_bufs[ins.buffer], self.grad_works[ins.buffer] = self.link.post_recv_grads()

    def _wait_recv_grad(self, ins):
        if ins.buffer not in self.grad_bufs:
            self.grad_bufs[ins.buffer], self.grad_works[ins.buffer] = self.link.post_recv_grads()
        StageLink.wait(self.grad_works.pop(ins.buffer, []))

    def _backward(self, ins):
        if self.is_last:
            loss = self.results.pop(ins.buffer)
            (loss / self.micro_batches).backward()
        else:
            outs = self.outputs.pop(ins.buffer)
            grads = self.grad_bufs.pop(ins.buffer)
            pairs = [(o, g) for o, g in zip(outs, grads) if g is not None and o.requires_grad]
            if pairs:
                torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])

    def _send_grad(self, ins):
        ins_t = self.inputs.pop(ins.buffer)
        self.link.send_grads([t.grad for t in ins_t], ins_t)

    _HANDLERS = {
        S.LoadMicroBatch: _load_micro_batch, S.PostRecvActivation: _post_recv_act,
        S.WaitRecvActivation: _wait_recv_act, S.ForwardPass: _forward, S.SendActivation: _send_act,
        S.PostRecvGrad: _post_recv_grad, S.WaitRecvGrad: _wait_recv_grad, S.BackwardPass: _backward,
        S.SendGrad: _send_grad,
    }

    def _run(self, sched: S.PipeSchedule):
        for step in sched:
            for ins in step:
                h = self._HANDLERS.get(type(ins))
                if h is None:
                    if isinstance(ins, S.RecvActivation):
                        self._post_recv_act(ins); self._wait_recv_act(ins)
                    elif isinstance(ins, S.RecvGrad):
                        self._post_recv_grad(ins); self._wait_recv_grad(ins)
                    continue  # ReduceGrads / OptimizerStep happen outside (engine hooks / user loop)
                h(self, ins)
        self.link.flush()
        if torch.is_grad_enabled():
            self._warm = True          # forward AND backward kernels are loaded now

    # ---- public ---------------------------------------------------------------------------------------------------
    def _prepare(self, kwargs: dict, output_fn, training: bool):
        from .microbatch import split_kwargs_into_chunks
        self._reset_state()
        self.output_fn = output_fn
        chunks = split_kwargs_into_chunks(kwargs, self.micro_batches)
        self.mb_kwargs = chunks
        if output_fn is not None:
            import inspect
            names = [n for n in inspect.signature(output_fn).parameters][1:]
            self.output_fn_kwargs = [{k: v for k, v in c.items() if k in names} for c in chunks]
            dev = self.device
            self.output_fn_kwargs = [{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
                                     for d in self.output_fn_kwargs]
        else:
            self.output_fn_kwargs = [{} for _ in chunks]

    def forward_backward(self, kwargs: dict, output_fn=None) -> torch.Tensor:
        self._prepare(kwargs, output_fn, True)
        sched = S.create_scheduler(self.algo, True, self.micro_batches, self.stages, self.stage)
        self._run(sched)
        return self._aggregate_loss()

    def forward(self, kwargs: dict, output_fn=None):
        self._prepare(kwargs, output_fn, False)
        with torch.no_grad():
            self._run(S.create_scheduler(self.algo, False, self.micro_batches, self.stages, self.stage))
        if not self.is_last:
            return None
        outs = [self.results[i] for i in range(self.micro_batches)]
        return outs[0] if len(outs) == 1 else outs

    def _aggregate_loss(self) -> torch.Tensor:
        """Mean loss over micro-batches, averaged over the data-parallel ranks, broadcast from the last stage
        (reference executor.py:283-321)."""
        if self.is_last:
            loss = (self.total_loss / self.micro_batches).reshape(1)
            g = self.mesh.get_data_proc_group()
            if g is not None:
                dist.all_reduce(loss, group=g)
                loss /= self.mesh.get_data_num()
        else:
            loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        if self.broadcast_loss and self.stages > 1:
            dist.broadcast(loss, src=self.mesh.stage_to_global(self.stages - 1), group=self.mesh.get_pp_proc_group())
        return loss[0]
