"""GPU negative samplers, bit-exact with hopwise's NumPy samplers.

Host-side mirror of (paths under /root/reference/hopwise/):
  sampler/sampler.py:294-357   KGSampler  (sample_by_entity_ids, used_ids, _uni_sampling)
  sampler/sampler.py:186-291   Sampler    (sample_by_user_ids, used_ids per phase)
  sampler/sampler.py:140-183   AbstractSampler.sample_by_key_ids
  utils/utils.py:203-220       init_seed -> np.random.seed(seed)

The reference draws from NumPy's *global* MT19937 stream, shared by the KG and the rec sampler
(KG draws first each step, knowledge_dataloader.py:137-145).  Here that stream is an explicit
``MTStream`` living on the device; both samplers advance it through the same kernel, so for
the same seed the produced ids equal the reference's, call after call.  ``MTStream`` converts
to and from ``np.random.get_state()`` tuples for hand-over with host code.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _abi


class MTStream:
    """A NumPy-legacy MT19937 stream held on the device: 624 key words, pos, status."""

    def __init__(self, seed=None, device="cuda", state=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("MTStream lives on a CUDA device; there is no CPU fallback")
        self.words = torch.zeros(626, dtype=torch.int32, device=self.device)
        if state is not None:
            self.set_state(state)
        elif seed is not None:
            self.seed(seed)
        else:
            self.seed(0)

    def seed(self, seed: int):
        """np.random.seed(seed) for an integer seed (legacy init_genrand)."""
        seed = int(seed)
        if not 0 <= seed <= 0xFFFFFFFF:
            raise ValueError("legacy integer seed must fit 32 bits")
        with _abi.on_device(self.device):
            _abi.check(_abi.lib().kge_mt19937_seed(self.words.data_ptr(), seed, _abi.stream_ptr()), "kge_mt19937_seed")

    def set_state(self, state):
        """Accepts np.random.get_state() tuples: ('MT19937', key[624], pos, ...)."""
        key = np.asarray(state[1], dtype=np.uint32)
        host = np.zeros(626, dtype=np.uint32)
        host[:624] = key
        host[624] = int(state[2])
        self.words.copy_(torch.from_numpy(host.view(np.int32)))

    def get_state(self):
        host = self.words.cpu().numpy().view(np.uint32)
        return ("MT19937", host[:624].copy(), int(host[624]), 0, 0.0)

    def exhausted(self) -> bool:
        """True when some call met a key whose forbidden list covers the whole range."""
        return bool(self.words[625].item())


def _as_device_ids(x, device):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=torch.int64).contiguous()
    return torch.as_tensor(np.asarray(x), dtype=torch.int64).to(device).contiguous()


def _as_host_ids(x):
    if torch.is_tensor(x):
        return x.detach().cpu().numpy().astype(np.int64, copy=False).reshape(-1)
    return np.asarray(x, dtype=np.int64).reshape(-1)


def build_used_csr(keys, values, n_keys: int, device):
    """CSR (offsets[n_keys+1], sorted unique values) of `used_ids[key] = set(values)`."""
    keys = _as_device_ids(keys, device)
    values = _as_device_ids(values, device)
    if keys.numel():
        span = int(values.max().item()) + 1
        packed = torch.unique(keys * span + values)  # sorted
        k2 = torch.div(packed, span, rounding_mode="floor")
        v2 = packed - k2 * span
        counts = torch.bincount(k2, minlength=n_keys)
    else:
        v2 = torch.zeros(0, dtype=torch.int64, device=device)
        counts = torch.zeros(n_keys, dtype=torch.int64, device=device)
    off = torch.zeros(n_keys + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts, 0, out=off[1:])
    return off, v2.contiguous()


def build_alias_table(candidates, alpha: float):
    """The reference's alias table (sampler.py:68-100) as three arrays in its key order:
    keys[n] int64 (first occurrence in `candidates`), prob[n] float64, alias[n] int64 (aliased key id, -1 = none).

    Host set-up, run once per sampler.  Every float is produced by the same IEEE-754 double operations, in the same
    order, as the reference's dict arithmetic (count / len, pow, the builtin sum, / total * n, and the running
    prob[large] - (1 - prob[small]) of the pairing loop), so `prob > p` decides identically; the two FIFO queues
    are deques instead of list.pop(0).
    """
    from collections import deque

    cand = np.asarray(candidates, dtype=np.int64).reshape(-1)
    if cand.size == 0:
        raise ValueError("popularity sampling needs at least one candidate")
    uniq, first, counts = np.unique(cand, return_index=True, return_counts=True)
    order = np.argsort(first, kind="stable")          # dict(Counter(...)) iterates in first-occurrence order
    keys = uniq[order]
    n_cand, n_keys = int(cand.size), int(keys.size)
    prob = [pow(c / n_cand, alpha) for c in counts[order].tolist()]
    total = sum(prob)
    large, small = deque(), deque()
    for i in range(n_keys):
        prob[i] = prob[i] / total * n_keys
        if prob[i] > 1:
            large.append(i)
        elif prob[i] < 1:
            small.append(i)
    alias = [-1] * n_keys
    while large and small:
        lq, sq = large.popleft(), small.popleft()
        alias[sq] = int(keys[lq])
        prob[lq] = prob[lq] - (1 - prob[sq])
        if prob[lq] < 1:
            small.append(lq)
        elif prob[lq] > 1:
            large.append(lq)
    return keys, np.asarray(prob, dtype=np.float64), np.asarray(alias, dtype=np.int64)


class _FilteredUniformSampler:
    """sample_by_key_ids (sampler.py:140-183) with uniform candidates in [1, value_num), or -- after
    ``set_popularity`` -- popularity-biased candidates from the alias table (sampler.py:102-116)."""

    def __init__(self, keys, values, n_keys: int, value_num: int, stream: MTStream, what: str):
        self.pop = None        # (keys, prob, alias) on the device
        self.stream = stream
        self.device = stream.device
        self.to_host = False   # hand the ids back as a CPU tensor (inside hopwise's own CPU loaders)
        self.value_num = int(value_num)
        self.n_keys = int(n_keys)
        self.used_off, self.used_vals = build_used_csr(keys, values, n_keys, self.device)
        counts = self.used_off[1:] - self.used_off[:-1]
        # the forbidden values that matter lie in [1, value_num)
        if counts.numel() and int(counts.max().item()) >= self.value_num - 1:
            in_range = (self.used_vals >= 1) & (self.used_vals < self.value_num)
            seg = torch.repeat_interleave(torch.arange(n_keys, device=self.device), counts)
            eff = torch.bincount(seg[in_range], minlength=n_keys)
            if int(eff.max().item()) >= self.value_num - 1:
                # sampler.py:241-249 / 329-336 raise the same way
                raise ValueError(f"Some {what} have interacted with all values, which we can not sample negatives for.")

    def set_popularity(self, table):
        """`table` = build_alias_table(...) (host arrays); shared between the phase copies of one sampler."""
        self.pop = tuple(torch.from_numpy(np.ascontiguousarray(x)).to(self.device) for x in table)

    def sample_by_key_ids(self, key_ids, num: int = 1) -> torch.Tensor:
        keys = _as_device_ids(key_ids, self.device)
        n = keys.numel()
        total = n * int(num)
        out = torch.empty(total, dtype=torch.int64, device=self.device)
        if total == 0:
            return out
        lib = _abi.lib()
        if self.pop is not None:
            pk, pp, pa = self.pop
            ws = torch.empty(max(lib.kge_sample_alias_workspace_bytes(total), 8), dtype=torch.uint8, device=self.device)
            with _abi.on_device(self.device):
                _abi.check(
                    lib.kge_sample_negatives_alias(
                        self.stream.words.data_ptr(), keys.data_ptr(), n, int(num), self.used_off.data_ptr(),
                        self.used_vals.data_ptr(), pk.numel(), pk.data_ptr(), pp.data_ptr(), pa.data_ptr(),
                        out.data_ptr(), ws.data_ptr(), _abi.stream_ptr(),
                    ),
                    "kge_sample_negatives_alias",
                )
            return out.cpu() if self.to_host else out
        ws = torch.empty(max(lib.kge_sample_workspace_bytes(total), 8), dtype=torch.uint8, device=self.device)
        with _abi.on_device(self.device):
            _abi.check(
                lib.kge_sample_negatives(
                    self.stream.words.data_ptr(), keys.data_ptr(), n, int(num), self.used_off.data_ptr(),
                    self.used_vals.data_ptr(), 1, self.value_num, out.data_ptr(), ws.data_ptr(), _abi.stream_ptr(),
                ),
                "kge_sample_negatives",
            )
        return out.cpu() if self.to_host else out


def _sets_from_csr(off, vals, n_keys):
    """np.ndarray (dtype=object) of Python sets, one per key -- the shape `used_ids` has in the reference
    (sampler.py:229-252, 321-336), which FullSortEvalDataLoader subtracts positives from
    (general_dataloader.py:204, 231-237).  Built on demand from the device CSR."""
    off = off.cpu().numpy()
    vals = vals.cpu().numpy()
    out = np.empty(n_keys, dtype=object)
    for k in range(n_keys):
        out[k] = set(vals[off[k]:off[k + 1]].tolist())
    return out


class KGSampler(_FilteredUniformSampler):
    """Tail corruption filtered by the head's true tails (any relation), sampler.py:294-357.

    ``KGSampler(dataset, distribution, alpha)`` reads ``dataset.head_entities / tail_entities / entity_num`` like
    the reference; arrays can be given directly instead.  ``distribution="popularity"`` draws candidates from the
    alias table over head+tail occurrences (sampler.py:68-116, 318-319).  ``used_ids`` is the reference's attribute (an array of
    sets indexed by head entity), materialised on first access.
    """

    def __init__(self, dataset=None, distribution="uniform", alpha=1.0, *, heads=None, tails=None, entity_num=None,
                 stream: MTStream | None = None, device="cuda", alias_table=None):
        if distribution not in ("uniform", "popularity"):
            raise NotImplementedError(f"The sampling distribution [{distribution}] has not been implemented.")
        if dataset is not None:
            heads, tails, entity_num = dataset.head_entities, dataset.tail_entities, dataset.entity_num
        self.distribution, self.alpha = distribution, alpha
        self.entity_num = int(entity_num)
        stream = stream if stream is not None else MTStream(device=device)
        super().__init__(heads, tails, self.entity_num, self.entity_num, stream, "head entities")
        self._used_sets = None
        if distribution == "popularity":
            # sampler.py:318-319: every head occurrence, then every tail occurrence
            cand = np.concatenate([_as_host_ids(heads), _as_host_ids(tails)])
            self.set_popularity(alias_table if alias_table is not None else build_alias_table(cand, alpha))

    @property
    def used_ids(self):
        if self._used_sets is None:
            self._used_sets = _sets_from_csr(self.used_off, self.used_vals, self.n_keys)
        return self._used_sets

    def sample_by_entity_ids(self, head_entity_ids, num: int = 1) -> torch.Tensor:
        """[len(heads) * num] int64 on the device, j-major: out[j*len + i] (sampler.py:338-357)."""
        return self.sample_by_key_ids(head_entity_ids, num)


class RecSampler:
    """Item negatives filtered by the user's interacted items: hopwise's ``Sampler`` (sampler.py:186-291).

    Two constructors:
      * ``RecSampler(phases, datasets)`` -- the reference signature: ``datasets[i].inter_feat`` holds the
        interactions of ``phases[i]``; the forbidden items of a phase are those of all phases up to it
        (cumulative, sampler.py:229-252).  ``set_phase(phase)`` returns the shallow copy bound to one phase, with
        ``.phase`` and ``.used_ids`` like the reference object the loaders hold (``train_data._sampler``,
        and the one FullSortEvalDataLoader reads, general_dataloader.py:204).
      * ``RecSampler(users, items, n_users, n_items)`` -- one phase given as id arrays (already bound).
    Every copy shares one ``MTStream``: NumPy's global generator in the reference.
    """

    def __init__(self, phases_or_users, datasets_or_items, n_users=None, n_items=None, distribution="uniform",
                 alpha=1.0, stream: MTStream | None = None, device="cuda", alias_table=None):
        if distribution not in ("uniform", "popularity"):
            raise NotImplementedError(f"The sampling distribution [{distribution}] has not been implemented.")
        self.distribution, self.alpha = distribution, alpha
        self.stream = stream if stream is not None else MTStream(device=device)
        self.device = self.stream.device
        self._impl, self._sets, self.phase = {}, {}, None
        by_phase = isinstance(phases_or_users, str) or (
            isinstance(phases_or_users, (list, tuple)) and all(isinstance(p, str) for p in phases_or_users))
        if by_phase:
            phases = [phases_or_users] if isinstance(phases_or_users, str) else list(phases_or_users)
            datasets = datasets_or_items if isinstance(datasets_or_items, (list, tuple)) else [datasets_or_items]
            if len(phases) != len(datasets):
                raise ValueError(f"Phases {phases} and datasets {datasets} should have the same length.")
            self.phases, self.datasets = phases, list(datasets)
            d0 = self.datasets[0]
            self.uid_field, self.iid_field = d0.uid_field, d0.iid_field
            self.user_num, self.item_num = int(d0.user_num), int(d0.item_num)
            users, items = [], []
            for phase, ds in zip(phases, self.datasets):   # cumulative: a phase forbids every earlier phase's items
                users.append(np.asarray(ds.inter_feat[self.uid_field]))
                items.append(np.asarray(ds.inter_feat[self.iid_field]))
                self._impl[phase] = _FilteredUniformSampler(np.concatenate(users), np.concatenate(items),
                                                            self.user_num, self.item_num, self.stream, "users")
        else:
            self.phases, self.datasets = ["train"], []
            self.user_num, self.item_num = int(n_users), int(n_items)
            self._impl["train"] = _FilteredUniformSampler(phases_or_users, datasets_or_items, self.user_num,
                                                          self.item_num, self.stream, "users")
            self.phase = "train"
            items = [_as_host_ids(datasets_or_items)]
        if distribution == "popularity":
            # sampler.py:220-224: the item column of every phase's dataset, concatenated -- one table for all phases
            # (the table depends on the ORDER of the interactions at this moment -- the key order is that of first
            # occurrence -- and hopwise's evaluation loaders sort their datasets by user later on,
            # general_dataloader.py:92: a sampler that replaces an existing one takes over its table, `alias_table`)
            table = alias_table if alias_table is not None else build_alias_table(
                np.concatenate([_as_host_ids(x) for x in items]), alpha)
            first = next(iter(self._impl.values()))
            first.set_popularity(table)
            for impl in self._impl.values():
                impl.pop = first.pop

    def set_phase(self, phase):
        """sampler.py:254-270: the copy of this sampler bound to `phase`."""
        if phase not in self.phases:
            raise ValueError(f"Phase [{phase}] not exist.")
        import copy

        bound = copy.copy(self)   # shallow: the CSRs, the set cache and the MT stream are shared
        bound.phase = phase
        return bound

    @property
    def used_ids(self):
        """Bound to a phase: array of sets indexed by user (what the eval loader reads); unbound: {phase: array}."""
        def sets(phase):
            if phase not in self._sets:
                impl = self._impl[phase]
                self._sets[phase] = _sets_from_csr(impl.used_off, impl.used_vals, impl.n_keys)
            return self._sets[phase]

        return sets(self.phase) if self.phase is not None else {p: sets(p) for p in self.phases}

    @property
    def used_off(self):
        return self._impl[self.phase or self.phases[0]].used_off

    @property
    def used_vals(self):
        return self._impl[self.phase or self.phases[0]].used_vals

    def sample_by_key_ids(self, key_ids, num: int = 1) -> torch.Tensor:
        if self.phase is None:
            raise RuntimeError("call set_phase() first (sampler.py:186-199)")
        return self._impl[self.phase].sample_by_key_ids(key_ids, num)

    def sample_by_user_ids(self, user_ids, item_ids=None, num: int = 1) -> torch.Tensor:
        """[len(users) * num] int64 on the device, j-major (sampler.py:272-291)."""
        return self.sample_by_key_ids(user_ids, num)


def _table_of(sampler):
    """(keys, prob, alias) arrays of a reference sampler's alias table (sampler.py:68-100: the dicts `prob` and
    `alias`, in key order), or None when it samples uniformly."""
    if getattr(sampler, "distribution", "uniform") != "popularity":
        return None
    keys = list(sampler.prob.keys())
    return (np.array([int(k) for k in keys], dtype=np.int64),
            np.array([float(sampler.prob[k]) for k in keys], dtype=np.float64),
            np.array([int(sampler.alias[k]) for k in keys], dtype=np.int64))


def install_device_samplers(train_data, device="cuda", to_host=True):
    """Swap the two CPU samplers of a hopwise ``KnowledgeBasedDataLoader`` for the device ones, in place:
    ``train_data.general_dataloader._sampler`` (rec negatives) and ``train_data.kg_dataloader._sampler`` (KG
    negatives, knowledge_dataloader.py:51,73).  Both continue NumPy's global MT19937 stream from where it stands
    (``np.random.get_state()``), sharing it exactly like the reference's samplers share the global generator, so the
    ids drawn are the ones the CPU samplers would have drawn.  hopwise's loaders assemble the batch on the host
    (``dataset.join`` indexes CPU feature tables with the sampled ids, abstract_dataloader.py:192-198), so with
    ``to_host`` (default) the ids come back as CPU tensors.  Returns the shared MTStream."""
    stream = MTStream(state=np.random.get_state(), device=device)
    gen, kg = train_data.general_dataloader, train_data.kg_dataloader
    old = gen._sampler
    rec = RecSampler(list(old.phases), list(old.datasets), distribution=old.distribution, alpha=old.alpha,
                     stream=stream, device=device, alias_table=_table_of(old)).set_phase(old.phase)
    old_kg = kg._sampler
    kgs = KGSampler(kg._dataset, distribution=getattr(old_kg, "distribution", "uniform"),
                    alpha=getattr(old_kg, "alpha", 1.0), stream=stream, device=device, alias_table=_table_of(old_kg))
    kgs.to_host = to_host
    for impl in rec._impl.values():
        impl.to_host = to_host
    gen._sampler = rec
    kg._sampler = kgs
    return stream
