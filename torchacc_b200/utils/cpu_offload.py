"""Activation offload to pinned host memory, overlapped with compute on the copy engines.

API parity: ``get_cpu_offload_context(num_offload_layers, num_prefetch_layers, num_offload_sync_layers, debug)``
returns ``(context, synchronizer)`` used as in reference torchacc/utils/cpu_offload.py:521-605:

    for layer in layers:
        with context:
            x = layer(x)
        x = synchronizer(x)

Design (own implementation; the reference adapts TransformerEngine's double-buffer handler):
a ``LayerOffloader`` numbers the ``with context`` regions ("groups") in forward order.  Inside a region a
saved-tensor hook replaces every activation that autograd saves by a ticket; its bytes are copied D2H on a
dedicated stream into a pooled page-locked buffer.  The synchronizer is an autograd identity whose
forward *commits* the group (records the D2H event; releases device memory of the group that is
``num_offload_sync_layers`` behind, once its copies are done) and whose backward *prefetches*: when backward
reaches group g, H2D copies for groups g-1 .. g-num_prefetch_layers are queued on a second stream so the
activations are resident again by the time they are needed.  Only groups ``< num_offload_layers`` are offloaded.
Views of one storage are offloaded once (keyed by storage pointer + offset + shape).
"""
from __future__ import annotations

import traceback
from typing import Dict, List, Optional

import torch

from .logger import logger


class _Ticket:
    __slots__ = ("group", "key", "meta", "host", "device", "d2h_done", "h2d_done", "refs")

    def __init__(self, group, key, meta):
        self.group, self.key, self.meta = group, key, meta
        self.host: Optional[torch.Tensor] = None
        self.device: Optional[torch.Tensor] = None
        self.d2h_done = None
        self.h2d_done = None
        self.refs = 0


class _HostPool:
    def __init__(self):
        self.free: Dict[tuple, List[torch.Tensor]] = {}

    def get(self, numel, dtype):
        lst = self.free.get((numel, dtype))
        if lst:
            return lst.pop()
        return torch.empty(numel, dtype=dtype, pin_memory=True)

    def put(self, t):
        self.free.setdefault((t.numel(), t.dtype), []).append(t)


class LayerOffloader:

    def __init__(self, num_offload_layers=1, num_prefetch_layers=1, num_offload_sync_layers=1.0, debug=False,
                 min_numel: int = 1024):
        self.num_offload = int(num_offload_layers)
        self.num_prefetch = max(int(num_prefetch_layers), 1)
        self.sync_window = max(int(round(num_offload_sync_layers)), 1)
        self.debug = debug
        self.min_numel = min_numel
        self.cur_group = 0
        self.in_region = False
        self.tickets: Dict[int, Dict[tuple, _Ticket]] = {}
        self.pool = _HostPool()
        self.d2h = None
        self.h2d = None
        self.offloaded_bytes = 0
        self.committed_groups = 0
        self.prefetched: set = set()

    # ---- helpers ------------------------------------------------------------------------------------------
    def _streams(self, device):
        if self.d2h is None:
            self.d2h = torch.cuda.Stream(device)
            self.h2d = torch.cuda.Stream(device)

    def _should_offload(self, t: torch.Tensor) -> bool:
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.numel() < self.min_numel:
            return False
        if self.cur_group >= self.num_offload or not self.in_region:
            return False
        if isinstance(t, torch.nn.Parameter):
            return False
        base = t._base
        if base is not None:                       # e.g. weight.T: a view of a leaf parameter -> keep on device
            return not base.is_leaf
        return (not t.is_leaf) or (not t.requires_grad)

    def begin_step_if_needed(self):
        if not torch.is_grad_enabled():
            return
        if self.cur_group == 0 and self.tickets:
            self.tickets = {}
            self.prefetched = set()

    # ---- saved-tensor hooks -----------------------------------------------------------------------------------
    def pack(self, t: torch.Tensor):
        if not self._should_offload(t):
            return t
        self._streams(t.device)
        key = (t.untyped_storage().data_ptr(), t.storage_offset(), tuple(t.shape), tuple(t.stride()), t.dtype)
        group = self.tickets.setdefault(self.cur_group, {})
        tk = group.get(key)
        if tk is None:
            tk = _Ticket(self.cur_group, key, (tuple(t.shape), t.dtype, t.device))
            src = t if t.is_contiguous() else t.contiguous()
            tk.host = self.pool.get(src.numel(), src.dtype)
            ready = torch.cuda.Event()
            ready.record()                                        # producer of `t` has been queued
            self.d2h.wait_event(ready)
            with torch.cuda.stream(self.d2h):
                tk.host.copy_(src.reshape(-1), non_blocking=True)
                tk.d2h_done = torch.cuda.Event()
                tk.d2h_done.record(self.d2h)
            src.record_stream(self.d2h)
            tk.device = src                                        # keep alive until the sync window passes
            group[key] = tk
            self.offloaded_bytes += src.numel() * src.element_size()
            if self.debug:
                logger.info("offload group %d: %s %s  (%.3f GiB total)\n%s", self.cur_group, tuple(t.shape), t.dtype,
                            self.offloaded_bytes / 2 ** 30, "".join(traceback.format_stack(limit=6)))
        tk.refs += 1
        return tk

    def unpack(self, obj):
        if not isinstance(obj, _Ticket):
            return obj
        tk = obj
        if tk.device is None:
            self._fetch(tk)
        if tk.h2d_done is not None:
            torch.cuda.current_stream().wait_event(tk.h2d_done)
        out = tk.device.view(tk.meta[0])
        tk.refs -= 1
        if tk.refs <= 0:
            if tk.host is not None:
                self.pool.put(tk.host)
                tk.host = None
            tk.device = None
        return out

    def _fetch(self, tk: _Ticket):
        shape, dtype, device = tk.meta
        with torch.cuda.stream(self.h2d):
            if tk.d2h_done is not None:
                self.h2d.wait_event(tk.d2h_done)
            tk.device = torch.empty(tk.host.numel(), dtype=dtype, device=device)
            tk.device.copy_(tk.host, non_blocking=True)
            tk.h2d_done = torch.cuda.Event()
            tk.h2d_done.record(self.h2d)
        tk.device.record_stream(torch.cuda.current_stream())

    # ---- group boundaries ---------------------------------------------------------------------------------------
    def commit_forward(self):
        """End of group ``cur_group`` in forward: drop device copies of the group that left the sync window."""
        g = self.cur_group
        old = g - self.sync_window + 1
        if 0 <= old < self.num_offload:
            for tk in self.tickets.get(old, {}).values():
                if tk.device is not None and tk.d2h_done is not None:
                    torch.cuda.current_stream().wait_event(tk.d2h_done)
                    tk.device = None
        self.cur_group = g + 1
        self.committed_groups = self.cur_group

    def finish_forward(self):
        for old in range(max(self.cur_group - self.sync_window + 1, 0), min(self.cur_group, self.num_offload)):
            for tk in self.tickets.get(old, {}).values():
                if tk.device is not None and tk.d2h_done is not None:
                    torch.cuda.current_stream().wait_event(tk.d2h_done)
                    tk.device = None

    def on_backward_reach(self, group: int):
        """Backward arrived at the boundary after ``group``: make group and the next ``num_prefetch`` earlier
        groups resident."""
        self.cur_group = 0   # next forward starts a new step
        for g in range(group, max(group - self.num_prefetch - 1, -1), -1):
            if g in self.prefetched or g >= self.num_offload:
                continue
            self.prefetched.add(g)
            for tk in self.tickets.get(g, {}).values():
                if tk.device is None and tk.host is not None:
                    self._fetch(tk)


class _GroupBoundary(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, offloader: LayerOffloader, group: int):
        ctx.offloader, ctx.group = offloader, group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ctx.offloader.on_backward_reach(ctx.group)
        return g, None, None


class _OffloadContext:
    """Re-enterable context manager installing the saved-tensor hooks for one group."""

    def __init__(self, offloader: LayerOffloader):
        self.offloader = offloader
        self._hooks = None

    def __enter__(self):
        o = self.offloader
        o.begin_step_if_needed()
        o.in_region = True
        self._hooks = torch.autograd.graph.saved_tensors_hooks(o.pack, o.unpack)
        self._hooks.__enter__()
        return self

    def __exit__(self, *exc):
        self._hooks.__exit__(*exc)
        self.offloader.in_region = False
        return False


def get_cpu_offload_context(num_offload_layers: int = 1, num_prefetch_layers: int = 1,
                            num_offload_sync_layers: float = 1.0, debug: bool = False):
    """Returns ``(context, synchronizer)`` -- see module docstring."""
    off = LayerOffloader(num_offload_layers, num_prefetch_layers, num_offload_sync_layers, debug)

    def synchronizer(outputs):
        if not torch.is_grad_enabled():
            return outputs
        group = off.cur_group
        off.commit_forward()
        from .utils import apply_to_tensors
        done = [False]

        def mark(t):
            if done[0] or not t.requires_grad:
                return t
            done[0] = True
            return _GroupBoundary.apply(t, off, group)

        return apply_to_tensors(mark, outputs)

    ctx = _OffloadContext(off)
    ctx.offloader = off
    return ctx, synchronizer
