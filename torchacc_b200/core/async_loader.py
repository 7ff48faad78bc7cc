"""AsyncLoader: background batch preparation + truly asynchronous host->device upload, with length bucketing.

Reference: torchacc/core/async_loader.py:14-207.  There the CUDA path is a synchronous ``tensor.to(device)`` that
ignores bucketing (``CUDALoader``, :141-156) and only the XLA path is threaded.
Here, for every device type:

* a producer thread pulls batches from the wrapped loader, pads the LAST dimension of every tensor up to the
  batch's bucket (explicit ``buckets`` or ``num_buckets`` uniform buckets up to ``max_length``) with the
  per-key pad value, and stages the result in page-locked host memory drawn from a small reusable pool;
* uploads are issued with ``non_blocking=True`` on a dedicated copy stream ``prefetch`` batches ahead of the
  consumer; the consumer's stream waits on a CUDA event per batch (no host synchronisation), and pinned staging
  buffers are recycled once their copy event has completed.
"""
from __future__ import annotations

import queue
import threading
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from ..utils.logger import logger
from ..utils.utils import apply_to_tensors

_DEFAULT_PAD = {"input_ids": 0, "attention_mask": 0, "labels": -100, "position_ids": 0}


def uniform_buckets(max_length: int, num_buckets: int = 8) -> List[int]:
    """``num_buckets`` evenly spaced bucket sizes ending at ``max_length``."""
    step = max(max_length // num_buckets, 1)
    b = [step * (i + 1) for i in range(num_buckets)]
    b[-1] = max(b[-1], max_length)
    return b


def closest_bucket(buckets: List[int], length: int) -> int:
    """Smallest bucket >= length; a longer-than-all sample opens a new bucket of exactly its length."""
    best = None
    for b in buckets:
        if b >= length and (best is None or b < best):
            best = b
    if best is None:
        buckets.append(length)
        best = length
    return best


class _PinnedPool:
    """Reusable page-locked staging buffers keyed by (shape, dtype)."""

    def __init__(self):
        self._free: Dict[tuple, List[torch.Tensor]] = {}
        self._busy: List[tuple] = []  # (event, key, buffer)

    def get(self, shape, dtype) -> torch.Tensor:
        self._reclaim()
        key = (tuple(shape), dtype)
        lst = self._free.get(key)
        if lst:
            return lst.pop()
        return torch.empty(shape, dtype=dtype, pin_memory=True)

    def give_back_after(self, event, buf):
        self._busy.append((event, (tuple(buf.shape), buf.dtype), buf))

    def _reclaim(self):
        still = []
        for ev, key, buf in self._busy:
            if ev.query():
                self._free.setdefault(key, []).append(buf)
            else:
                still.append((ev, key, buf))
        self._busy = still


class _Iterator:
    _END = object()

    def __init__(self, parent: "AsyncLoader"):
        self.p = parent
        self.device = parent._device if isinstance(parent._device, torch.device) else torch.device("cuda", parent._device)
        self.cuda = self.device.type == "cuda"
        self.q: "queue.Queue" = queue.Queue(maxsize=max(parent.prefetch, 1))
        self.copy_stream = torch.cuda.Stream(self.device) if self.cuda else None
        self.pool = _PinnedPool() if (self.cuda and parent.pin_memory) else None
        self.err = None
        self._stop = threading.Event()
        self.thread = threading.Thread(target=self._produce, daemon=True)
        self.thread.start()

    # ---- producer thread --------------------------------------------------------------------------------
    def _pad(self, batch):
        p = self.p
        if p.buckets is None:
            return batch
        lengths = []
        apply_to_tensors(lambda t: lengths.append(t.shape[-1]) if t.dim() >= 2 else None, batch)
        if not lengths:
            return batch
        target = closest_bucket(p.buckets, max(lengths))
        logger.debug("AsyncLoader: batch length %d -> bucket %d", max(lengths), target)

        def pad_one(t, key=None):
            if t.dim() < 2 or t.shape[-1] >= target:
                return t
            val = p.pad_value_dict.get(key, 0) if key is not None else 0
            return F.pad(t, (0, target - t.shape[-1]), value=val)

        if isinstance(batch, dict):
            return type(batch)({k: (pad_one(v, k) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()})
        return apply_to_tensors(pad_one, batch)

    def _upload(self, batch):
        if not self.cuda:
            return batch, None
        bufs = []

        def to_dev(t):
            if t.is_cuda:
                return t
            src = t
            if self.pool is not None and not t.is_pinned():
                src = self.pool.get(t.shape, t.dtype)
                src.copy_(t)
                bufs.append(src)
            return src.to(self.device, non_blocking=True)

        with torch.cuda.stream(self.copy_stream):
            out = apply_to_tensors(to_dev, batch)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        for b in bufs:
            self.pool.give_back_after(ev, b)
        return out, ev

    def _produce(self):
        try:
            if self.cuda:
                torch.cuda.set_device(self.device)
            for batch in self.p._loader:
                if self._stop.is_set():
                    return
                item = self._upload(self._pad(batch))
                self.q.put(item)
        except BaseException as e:  # surfaced in the consumer
            self.err = e
        finally:
            self.q.put(self._END)

    # ---- consumer -----------------------------------------------------------------------------------------
    def __iter__(self):
        return self

    def __next__(self):
        item = self.q.get()
        if item is self._END:
            if self.err is not None:
                raise self.err
            raise StopIteration
        batch, ev = item
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
            # tensors were allocated on the copy stream: tell the allocator the compute stream uses them too
            apply_to_tensors(lambda t: t.record_stream(torch.cuda.current_stream(self.device)) if t.is_cuda else t, batch)
        return batch

    def close(self):
        self._stop.set()

    def __del__(self):
        self._stop.set()


class AsyncLoader:
    """Wraps a DataLoader (any iterable of tensors / dicts / tuples) -- see module docstring.

    Args follow the reference (async_loader.py:159-192): ``loader, device, buckets=None, max_length=None,
    num_buckets=8, pad_value_dict=None`` plus ``prefetch`` (batches uploaded ahead) and ``pin_memory``.
    """

    def __init__(self, loader, device, buckets: Optional[List[int]] = None, max_length: Optional[int] = None,
                 num_buckets: Optional[int] = 8, pad_value_dict: Optional[Dict[str, int]] = None, prefetch: int = 2,
                 pin_memory: bool = True, **kwargs):
        self._loader = loader
        self._device = device
        if buckets is not None:
            self.buckets = sorted(buckets)
        elif max_length is not None:
            self.buckets = uniform_buckets(max_length, num_buckets or 8)
        else:
            self.buckets = None
        self.max_length, self.num_buckets = max_length, num_buckets
        self.pad_value_dict = dict(_DEFAULT_PAD)
        if pad_value_dict:
            self.pad_value_dict.update(pad_value_dict)
        self.prefetch, self.pin_memory = prefetch, pin_memory
        self._kwargs = kwargs

    def __iter__(self):
        return _Iterator(self)

    def __len__(self):
        return len(self._loader)

    @property
    def dataset(self):
        return getattr(self._loader, "dataset", None)

    @property
    def batch_size(self):
        return getattr(self._loader, "batch_size", None)
