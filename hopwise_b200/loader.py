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

import ctypes as C
from collections import deque

import numpy as np
import torch

from . import _abi


def _widen(src: torch.Tensor, dst: torch.Tensor, stream_ptr):
    _abi.check(_abi.lib().kge_widen_ids_i32(src.data_ptr(), dst.data_ptr(), src.numel(), stream_ptr), "kge_widen_ids_i32")


class PackedBatch:
    """Named id vectors laid out back to back in one (pinned) host tensor: int64, or int32 (`pack_batch(narrow=True)`)
    to halve the bytes that cross PCIe -- the device side widens them (`kge_widen_ids_i32`)."""

    def __init__(self, base: torch.Tensor, slices: dict):
        self.base, self.slices = base, dict(slices)

    @property
    def narrow(self) -> bool:
        return self.base.dtype == torch.int32

    def views(self, base=None):
        base = self.base if base is None else base
        return {k: base[o : o + n] for k, (o, n) in self.slices.items()}

    def to(self, device):
        dev = self.base.to(device, non_blocking=True)
        if self.narrow:
            wide = torch.empty(dev.numel(), dtype=torch.int64, device=dev.device)
            with _abi.on_device(dev.device):
                _widen(dev, wide, _abi.stream_ptr())
            dev = wide
        return self.views(dev)


def pack_batch(batch: dict, pin: bool = True, narrow: bool = False) -> PackedBatch:
    """Copy a dict of id vectors into one contiguous buffer (done once, by the loader): int64, or int32 with
    `narrow` (ids must be below 2^31)."""
    arrs = {k: torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v, dtype=torch.int64).reshape(-1)
            for k, v in batch.items()}
    if narrow and any(a.numel() and (int(a.max()) >= 2 ** 31 or int(a.min()) < -(2 ** 31)) for a in arrs.values()):
        raise ValueError("pack_batch(narrow=True): an id does not fit 32 bits")
    base = torch.empty((sum(a.numel() for a in arrs.values()) + 3) // 4 * 4, dtype=torch.int32 if narrow else torch.int64)
    if pin:
        base = base.pin_memory()
    slices, off = {}, 0
    for k, a in arrs.items():
        base[off : off + a.numel()].copy_(a)
        slices[k] = (off, a.numel())
        off += a.numel()
    base[off:] = 0   # (padding to a multiple of four ids: the widening kernel moves 16-byte quads)
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
        """Issue the copies of one batch on the copy stream (which must be current for foreign batch types)."""
        host = _host_tensors(batch)
        if host is None:   # foreign batch type: its own .to(), allocator-managed lifetime
            with torch.cuda.stream(self._stream):
                db = batch.to(self.device)
            inner = getattr(db, "interaction", None)
            for t in (inner.values() if isinstance(inner, dict) else []):
                if torch.is_tensor(t):
                    t.record_stream(torch.cuda.current_stream(self.device))
            return db
        from . import _abi

        lib, sptr = _abi.lib(), self._stream.cuda_stream
        # the raw copies below are invisible to torch's pinned-memory allocator: the slot keeps the host tensors
        # alive until it is staged again, which only happens after the consumer is done with this batch
        slot["__host__"] = host
        out = {}
        for k, v in host.items():
            buf = slot.get("dev:" + k)
            if buf is None or buf.dtype != v.dtype or buf.numel() < v.numel():
                with torch.cuda.stream(self._stream):   # the slot's memory belongs to the copy stream's pool
                    buf = torch.empty(v.numel(), dtype=v.dtype, device=self.device)
                slot["dev:" + k] = buf
            dst = buf[: v.numel()].view(v.shape)
            if v.is_contiguous() and not v.is_cuda:
                # one cudaMemcpyAsync on the copy stream (no stream switch on the host: ~20 us less per step)
                _abi.check(lib.kge_copy_h2d_async(dst.data_ptr(), v.data_ptr(), v.numel() * v.element_size(), sptr),
                           "kge_copy_h2d_async")
            else:
                with torch.cuda.stream(self._stream):
                    dst.copy_(v, non_blocking=True)
            out[k] = dst
        if isinstance(batch, PackedBatch):
            base = out["__base__"]
            if batch.narrow:   # widen on the copy stream too: the consumer sees int64 ids, as from any loader
                wide = slot.get("dev:__wide__")
                if wide is None or wide.numel() < base.numel():
                    with torch.cuda.stream(self._stream):
                        wide = torch.empty(base.numel(), dtype=torch.int64, device=self.device)
                    slot["dev:__wide__"] = wide
                with _abi.on_device(self.device):
                    _widen(base, wide[: base.numel()], sptr)
                base = wide[: base.numel()]
            return batch.views(base)
        return out

    def __iter__(self):
        it = iter(self.batches)
        queue = deque()
        n_slots = len(self._slots)
        # per slot: the copy's completion event and the consumer-side event after which the slot may be
        # overwritten (created once, re-recorded every use)
        if not hasattr(self, "_copied"):
            self._copied = [torch.cuda.Event() for _ in range(n_slots)]
            self._released = [torch.cuda.Event() for _ in range(n_slots)]
        used = [False] * n_slots
        nxt = 0

        def issue():
            nonlocal nxt
            b = next(it, None)
            if b is None:
                return
            idx = nxt % n_slots
            nxt += 1
            if used[idx]:
                self._stream.wait_event(self._released[idx])
                # the slot's previous host buffer is only kept alive by the slot (the raw cudaMemcpyAsync is invisible
                # to torch's pinned allocator): its DMA must have run before the reference is dropped -- in a loop
                # that never syncs (FusedKGTrainer's epoch) the GPU may lag the host by several batches
                self._copied[idx].synchronize()
            db = self._stage(self._slots[idx], b)
            self._copied[idx].record(self._stream)
            queue.append((db, idx))

        for _ in range(self.depth):
            issue()
        prev = None
        while queue:
            db, idx = queue.popleft()
            cur = torch.cuda.current_stream(self.device)
            if prev is not None:   # everything the consumer enqueued for the previous batch is behind this event
                self._released[prev].record(cur)
                used[prev] = True
            cur.wait_event(self._copied[idx])
            issue()
            prev = idx
            yield db


# ---- device-resident training loader ---------------------------------------------------------------------
class EpochOrder:
    """The index stream of one reference DataLoader, draw for draw.

    abstract_dataloader.py:47-73 builds ``DataLoader(list(range(n)), batch_size=step, shuffle=True,
    generator=torch.Generator().manual_seed(seed))``.  Every ``__iter__`` of it draws one int64
    ``random_()`` (the iterator's base seed) from that generator; the first ``next`` draws the epoch's
    ``randperm(n)``; and an iterator that is advanced past its last batch draws one more, discarded,
    ``randperm(n)`` (RandomSampler's trailing ``[: num_samples % n]`` slice).  Unshuffled loaders still
    draw the base seed.  The generator lives on the CPU, exactly like the reference's.

    With ``world > 1`` it is the multi-GPU loader of abstract_dataloader.py:59-64 instead: a
    ``DistributedSampler(list(range(n)), shuffle=shuffle, drop_last=False)`` picks the indices -- the epoch's
    permutation comes from ``Generator().manual_seed(0 + epoch)`` (the sampler's own seed 0, the epoch set through
    ``set_epoch``, trainer.py:240-241 and knowledge_dataloader.py:170-172), is padded by wrapping around to a
    multiple of ``world`` and strided ``rank::world`` -- and the batch is ``max(1, step // world)`` rows per rank.
    The loader's generator still draws its base seed per ``__iter__`` but no longer decides the order.
    """

    def __init__(self, n: int, step: int, seed: int, shuffle: bool = True, rank: int = 0, world: int = 1):
        self.n, self.shuffle = int(n), bool(shuffle)
        self.rank, self.world = int(rank), int(world)
        if not 0 <= self.rank < self.world:
            raise ValueError("rank outside [0, world)")
        self.step = int(step) if self.world == 1 else max(1, int(step) // self.world)
        self.generator = torch.Generator()
        self.generator.manual_seed(int(seed))
        self.epoch = 0
        self._perm = None
        self._pos = 0
        self._live = False
        self.mirror = None       # a CUDA device: the epoch's index vector is uploaded once, batches are views of it
        self._perm_dev = None

    @property
    def n_local(self) -> int:
        """Indices this rank visits per epoch (ceil(n / world): the tail is padded with repeats)."""
        return self.n if self.world == 1 else (self.n + self.world - 1) // self.world

    def __len__(self):
        return (self.n_local + self.step - 1) // self.step

    def set_epoch(self, epoch: int):
        """DistributedSampler.set_epoch (only matters when world > 1)."""
        self.epoch = int(epoch)

    def _distributed_indices(self):
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(0 + self.epoch)            # torch.utils.data.DistributedSampler: self.seed + self.epoch
            idx = torch.randperm(self.n, generator=g)
        else:
            idx = torch.arange(self.n)
        total = self.n_local * self.world
        pad = total - self.n
        if pad > 0:                                   # drop_last=False: wrap around (repeat if n < pad)
            reps = (pad + self.n - 1) // self.n
            idx = torch.cat([idx, idx.repeat(reps)[:pad]])
        return idx[self.rank : total : self.world]

    def start(self):
        """DataLoader.__iter__()."""
        torch.empty((), dtype=torch.int64).random_(generator=self.generator)
        self._perm, self._pos, self._live = None, 0, True

    def next_indices(self):
        """Indices of the next batch (int64 CPU tensor), or None once the iterator is exhausted."""
        if not self._live:
            raise RuntimeError("EpochOrder.start() must be called first")
        if self._perm is None:
            if self.world > 1:
                self._perm = self._distributed_indices()
            else:
                self._perm = torch.randperm(self.n, generator=self.generator) if self.shuffle else torch.arange(self.n)
        if self._pos >= self._perm.numel():
            if self.shuffle and self.world == 1:
                torch.randperm(self.n, generator=self.generator)   # the discarded trailing draw (RandomSampler)
            self._live = False
            return None
        idx = self._perm[self._pos : self._pos + self.step]
        self._pos += self.step
        return idx

    def next_indices_device(self):
        """next_indices() as a view of the device copy of the epoch's index vector (one upload per epoch instead of
        one per batch); None once exhausted."""
        had = self._perm is not None
        idx = self.next_indices()
        if idx is None:
            return None
        if not had or self._perm_dev is None:
            self._perm_dev = self._perm.to(self.mirror)
        pos = self._pos - self.step
        return self._perm_dev[pos : pos + idx.numel()]


class DeviceKGLoader:
    """KnowledgeBasedDataLoader in RSKG mode with the data, the gathers and both samplers on the GPU.

    Mirrors knowledge_dataloader.py:78-178 + general_dataloader.py:66-70 + abstract_dataloader.py:165-198:
    per step the KG batch is drawn first (``kg_feat[index]`` -> heads -> ``sample_by_entity_ids``), then the
    recommendation batch (``inter_feat[index]`` -> ``sample_by_user_ids``), and the two dicts are merged;
    an epoch is as long as the recommendation loader; the KG iterator restarts with a fresh permutation
    every epoch and wraps around if it runs out.  The batch order comes from the reference's CPU torch
    generators (``EpochOrder``); the interaction / triple arrays live on the device, so per step only the
    two index vectors (2 x batch x 8 B) cross PCIe instead of the 7 id vectors, and the negatives are
    drawn by the MT19937 kernel from the stream the two samplers share (KG draws before rec draws, H3).

    ``sample(kind, ids, num)`` is the sampling hook (tests substitute the CPU oracle for it).
    """

    KEYS = ("user_id", "item_id", "neg_item_id", "head_id", "relation_id", "tail_id", "neg_tail_id")

    def __init__(self, inter_user, inter_item, kg_head, kg_rel, kg_tail, rec_sampler, kg_sampler, batch_size: int,
                 seed: int, device="cuda", shuffle: bool = True, neg_sample_num: int = 1, gather=None,
                 rank: int = 0, world: int = 1, dynamic: bool = False, candidate_num: int = 0):
        self.device = torch.device(device)
        as_dev = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.int64).to(self.device)  # noqa: E731
        self.inter_user, self.inter_item = as_dev(inter_user), as_dev(inter_item)
        self.kg_head, self.kg_rel, self.kg_tail = as_dev(kg_head), as_dev(kg_rel), as_dev(kg_tail)
        self.rec_sampler, self.kg_sampler = rec_sampler, kg_sampler
        # (several GPUs: each rank walks its DistributedSampler shard with batch_size // world rows per step)
        self.rec_order = EpochOrder(self.inter_user.numel(), batch_size, seed, shuffle, rank, world)
        self.kg_order = EpochOrder(self.kg_head.numel(), batch_size, seed, True, rank, world)   # "must shuffle"
        self.neg_sample_num = int(neg_sample_num)
        # train_neg_sample_args["dynamic"] / ["candidate_num"] (abstract_dataloader.py:166-183): the negative of a
        # row is the best-scoring of `candidate_num` filtered candidates under the current weights
        self.dynamic, self.candidate_num = bool(dynamic), int(candidate_num)
        if self.dynamic and self.candidate_num < 1:
            raise ValueError("dynamic negative sampling needs candidate_num >= 1")
        self.model = None
        # `gather(table, idx)` is the hook the CPU host-logic tests drive the loader through; on the GPU the id
        # columns of a half are gathered by one kge_gather_columns launch
        self._gather = gather
        if gather is None and self.device.type != "cuda":
            raise RuntimeError("DeviceKGLoader gathers on a CUDA device; there is no CPU fallback")
        if gather is None:
            self.rec_order.mirror = self.kg_order.mirror = self.device
        self._col_ptrs = {}

    def get_model(self, model):
        """abstract_dataloader.py:214-215: the model dynamic negative sampling scores its candidates with."""
        self.model = model

    def _rec_negatives(self, user, item):
        num = self.neg_sample_num
        if not self.dynamic:
            return self.rec_sampler.sample_by_user_ids(user, item, num)
        if self.model is None:
            raise RuntimeError("dynamic negative sampling: call get_model(model) first (trainer.py:306)")
        cn = self.candidate_num
        cand = self.rec_sampler.sample_by_user_ids(user, item, num * cn)   # j-major: [cn, num * len(user)]
        with torch.no_grad():
            scores = self.model.predict({"user_id": user.repeat(num * cn), "item_id": cand}).reshape(cn, -1)
        best = torch.max(scores, dim=0)[1]
        return cand.reshape(cn, -1).gather(0, best.unsqueeze(0)).view(-1)

    def __len__(self):
        return len(self.rec_order)

    def set_epoch(self, epoch: int):
        """What the trainer does through ``train_data.sampler.set_epoch`` / ``knowledge_shuffle`` per epoch."""
        self.rec_order.set_epoch(epoch)
        self.kg_order.set_epoch(epoch)

    def _index(self, idx):
        if self.device.type == "cuda":
            idx = idx.pin_memory()
        return idx.to(self.device, non_blocking=True)

    def _take(self, order, columns):
        """The next batch of `order`: the id columns gathered by its index vector (interaction.py:130-139), or None
        when the order is exhausted."""
        if self._gather is not None:
            idx = order.next_indices()
            if idx is None:
                return None
            idx = self._index(idx)
            return [self._gather(c, idx) for c in columns]
        idx = order.next_indices_device()
        if idx is None:
            return None
        n, nc = idx.numel(), len(columns)
        out = torch.empty((nc, n), dtype=torch.int64, device=self.device)
        key = id(columns[0])
        src = self._col_ptrs.get(key)
        if src is None:
            src = self._col_ptrs[key] = (C.c_void_p * nc)(*[c.data_ptr() for c in columns])
        base = out.data_ptr()
        dst = (C.c_void_p * nc)(*[base + 8 * n * c for c in range(nc)])
        with _abi.on_device(self.device):
            _abi.check(
                _abi.lib().kge_gather_columns(src, nc, columns[0].numel(), idx.data_ptr(), n, dst, None,
                                              _abi.stream_ptr()),
                "kge_gather_columns",
            )
        return list(out.unbind(0))

    def _fast_path(self):
        """True when a whole batch can be assembled by one kge_assemble_batch call: CUDA gathers, both samplers this
        package's uniform ones on one stream, handing ids over on the device, static negatives."""
        from .sampler import KGSampler, RecSampler

        kg, rec = self.kg_sampler, self.rec_sampler
        if self._gather is not None or self.dynamic or not isinstance(kg, KGSampler) or not isinstance(rec, RecSampler):
            return False
        if rec.phase is None or kg.pop is not None or kg.to_host:
            return False
        impl = rec._impl[rec.phase]
        return impl.pop is None and not impl.to_host and impl.stream is kg.stream and kg.device == self.device

    def _assembled(self):
        """The epoch's batches through kge_assemble_batch: per step two index views, one output buffer, one call."""
        lib = _abi.lib()
        kg, impl = self.kg_sampler, self.rec_sampler._impl[self.rec_sampler.phase]
        num = self.neg_sample_num
        ws, ws_bytes = None, -1
        self.kg_order.start()
        self.rec_order.start()
        while True:
            kidx = self.kg_order.next_indices_device()
            if kidx is None:
                self.kg_order.start()
                kidx = self.kg_order.next_indices_device()
            ridx = self.rec_order.next_indices_device()
            n_kg = kidx.numel()
            n_rec = 0 if ridx is None else ridx.numel()
            need = lib.kge_assemble_batch_workspace_bytes(n_kg, n_rec, num)
            if need > ws_bytes:
                ws, ws_bytes = torch.empty(max(need, 8), dtype=torch.uint8, device=self.device), need
            out = torch.empty(4 * n_kg + (2 + num) * n_rec, dtype=torch.int64, device=self.device)
            with _abi.on_device(self.device):
                _abi.check(
                    lib.kge_assemble_batch(
                        kg.stream.words.data_ptr(), self.kg_head.data_ptr(), self.kg_rel.data_ptr(),
                        self.kg_tail.data_ptr(), self.kg_head.numel(), kidx.data_ptr(), n_kg, kg.used_off.data_ptr(),
                        kg.used_vals.data_ptr(), kg.value_num, self.inter_user.data_ptr(), self.inter_item.data_ptr(),
                        self.inter_user.numel(), None if ridx is None else ridx.data_ptr(), n_rec, num,
                        impl.used_off.data_ptr(), impl.used_vals.data_ptr(), impl.value_num, out.data_ptr(),
                        ws.data_ptr(), _abi.stream_ptr()),
                    "kge_assemble_batch",
                )
            if ridx is None:     # (the KG draw before the recommendation side is found exhausted: like the reference)
                return
            head, rel, tail, neg_tail, user, item, neg_item = out.split([n_kg] * 4 + [n_rec, n_rec, num * n_rec])
            if num > 1:
                user, item = user.repeat(num), item.repeat(num)
            yield {"head_id": head, "relation_id": rel, "tail_id": tail, "neg_tail_id": neg_tail, "user_id": user,
                   "item_id": item, "neg_item_id": neg_item}

    def __iter__(self):
        if self._fast_path():
            yield from self._assembled()
            return
        self.kg_order.start()    # knowledge_dataloader.py:131-135: kg iterator first, then the general one
        self.rec_order.start()
        while True:
            kg_cols = (self.kg_head, self.kg_rel, self.kg_tail)
            got = self._take(self.kg_order, kg_cols)
            if got is None:      # :137-142 wraps a KG loader that ran out
                self.kg_order.start()
                got = self._take(self.kg_order, kg_cols)
            head, rel, tail = got
            batch = {"head_id": head, "relation_id": rel, "tail_id": tail,
                     "neg_tail_id": self.kg_sampler.sample_by_entity_ids(head, 1)}   # knowledge_dataloader.py:51
            got = self._take(self.rec_order, (self.inter_user, self.inter_item))
            if got is None:
                return
            user, item = got
            batch["neg_item_id"] = self._rec_negatives(user, item)
            if self.neg_sample_num > 1:   # abstract_dataloader.py:191: the positives repeat once per negative
                user, item = user.repeat(self.neg_sample_num), item.repeat(self.neg_sample_num)
            batch["user_id"], batch["item_id"] = user, item
            yield batch
