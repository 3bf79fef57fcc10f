"""Host -> device staging of training batches, one batch ahead on a copy stream.

The reference trainer moves every batch with a blocking ``interaction.to(self.device)`` right before
``calculate_loss`` (trainer/trainer.py:250-256), so the H2D copy of the 7 id vectors sits on the critical
path of every step.  ``DevicePrefetcher`` wraps any iterable of batches (dicts of tensors / numpy arrays,
or objects with ``.to(device)`` such as hopwise's ``Interaction``) and issues the copy of batch i+1 on a
side stream while batch i trains; the consumer stream only waits on the copy's event.  Pinned host
tensors make the copies truly asynchronous.  Nothing else changes: the batches arrive in the same order
with the same contents.

A loader that assembles a batch in ONE pinned buffer can hand it over as ``pack_batch(...)``: the
prefetcher then issues a single copy per step and rebuilds the named views on the device (seven
separate copies cost ~0.1 ms of host time per step, which a trainer that reads ``loss.item()`` every step
cannot hide).
"""

from __future__ import annotations

from collections import deque

import numpy as np
import torch


class PackedBatch:
    """Named int64 id vectors laid out back to back in one (pinned) host tensor."""

    def __init__(self, base: torch.Tensor, slices: dict):
        self.base, self.slices = base, dict(slices)

    def views(self, base=None):
        base = self.base if base is None else base
        return {k: base[o : o + n] for k, (o, n) in self.slices.items()}

    def to(self, device):
        return self.views(self.base.to(device, non_blocking=True))


def pack_batch(batch: dict, pin: bool = True) -> PackedBatch:
    """Copy a dict of id vectors into one contiguous int64 buffer (done once, by the loader)."""
    arrs = {k: torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v, dtype=torch.int64).reshape(-1)
            for k, v in batch.items()}
    base = torch.empty(sum(a.numel() for a in arrs.values()), dtype=torch.int64)
    if pin:
        base = base.pin_memory()
    slices, off = {}, 0
    for k, a in arrs.items():
        base[off : off + a.numel()].copy_(a)
        slices[k] = (off, a.numel())
        off += a.numel()
    return PackedBatch(base, slices)


def _host_tensors(batch):
    """name -> host tensor for the batch kinds the staging ring understands, else None."""
    if isinstance(batch, PackedBatch):
        return {"__base__": batch.base}
    if isinstance(batch, dict):
        out = {}
        for k, v in batch.items():
            if isinstance(v, np.ndarray):
                v = torch.from_numpy(v)
            if not torch.is_tensor(v):
                return None
            out[k] = v
        return out
    return None


_COPY_STREAMS: dict = {}


class DevicePrefetcher:
    """Iterate `batches` with their tensors already on `device`.

    `depth` batches are in flight.  Device staging buffers are preallocated and recycled (a ring of
    depth + 1 slots), so a step makes no allocator call; a slot is overwritten only after the consumer
    stream has passed the work of the batch that used it.  A yielded batch is valid until the second
    batch after it is requested (as with any double-buffered loader).  Batches that are neither dicts of
    tensors nor `PackedBatch` (e.g. hopwise's Interaction) are moved with their own ``.to(device)`` on
    the copy stream instead.  Iterate the same instance every epoch (assign ``.batches`` to change the
    source): the staging buffers are kept.
    """

    def __init__(self, batches, device, depth: int = 2):
        self.batches, self.device, self.depth = batches, torch.device(device), max(1, int(depth))
        if self.device.type != "cuda":
            raise ValueError("DevicePrefetcher stages batches onto a CUDA device")
        # one copy stream per device for the whole process: a new stream would start with an empty
        # allocator pool and pay cudaMalloc for its staging buffers again
        self._stream = _COPY_STREAMS.get(self.device)
        if self._stream is None:
            self._stream = _COPY_STREAMS[self.device] = torch.cuda.Stream(self.device)
        self._slots = [dict() for _ in range(self.depth + 1)]

    def __len__(self):
        return len(self.batches)

    def _stage(self, slot: dict, batch):
        host = _host_tensors(batch)
        if host is None:   # foreign batch type: its own .to(), allocator-managed lifetime
            db = batch.to(self.device)
            inner = getattr(db, "interaction", None)
            for t in (inner.values() if isinstance(inner, dict) else []):
                if torch.is_tensor(t):
                    t.record_stream(torch.cuda.current_stream(self.device))
            return db
        out = {}
        for k, v in host.items():
            buf = slot.get(k)
            if buf is None or buf.dtype != v.dtype or buf.numel() < v.numel():
                buf = torch.empty(v.numel(), dtype=v.dtype, device=self.device)
                slot[k] = buf
            dst = buf[: v.numel()].view(v.shape)
            dst.copy_(v, non_blocking=True)
            out[k] = dst
        if isinstance(batch, PackedBatch):
            return batch.views(out["__base__"])
        return out

    def __iter__(self):
        it = iter(self.batches)
        queue = deque()
        n_slots = len(self._slots)
        released = [None] * n_slots   # consumer-side event after which a slot may be overwritten
        nxt = 0

        def issue():
            nonlocal nxt
            b = next(it, None)
            if b is None:
                return
            idx = nxt % n_slots
            nxt += 1
            with torch.cuda.stream(self._stream):
                if released[idx] is not None:
                    self._stream.wait_event(released[idx])
                db = self._stage(self._slots[idx], b)
                ev = torch.cuda.Event()
                ev.record(self._stream)
            queue.append((db, ev, idx))

        for _ in range(self.depth):
            issue()
        prev = None
        while queue:
            db, ev, idx = queue.popleft()
            cur = torch.cuda.current_stream(self.device)
            if prev is not None:   # everything the consumer enqueued for the previous batch is behind this event
                released[prev] = torch.cuda.Event()
                released[prev].record(cur)
            cur.wait_event(ev)
            issue()
            prev = idx
            yield db
