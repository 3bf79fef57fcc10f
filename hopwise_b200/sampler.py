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
        with torch.cuda.device(self.device):
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


class _FilteredUniformSampler:
    """sample_by_key_ids with uniform candidates in [1, value_num) (sampler.py:140-183)."""

    def __init__(self, keys, values, n_keys: int, value_num: int, stream: MTStream, what: str):
        self.stream = stream
        self.device = stream.device
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

    def sample_by_key_ids(self, key_ids, num: int = 1) -> torch.Tensor:
        keys = _as_device_ids(key_ids, self.device)
        n = keys.numel()
        total = n * int(num)
        out = torch.empty(total, dtype=torch.int64, device=self.device)
        if total == 0:
            return out
        lib = _abi.lib()
        ws = torch.empty(max(lib.kge_sample_workspace_bytes(total), 8), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _abi.check(
                lib.kge_sample_negatives(
                    self.stream.words.data_ptr(), keys.data_ptr(), n, int(num), self.used_off.data_ptr(),
                    self.used_vals.data_ptr(), 1, self.value_num, out.data_ptr(), ws.data_ptr(), _abi.stream_ptr(),
                ),
                "kge_sample_negatives",
            )
        return out


class KGSampler(_FilteredUniformSampler):
    """Tail corruption filtered by the head's true tails (any relation), sampler.py:294-357.

    ``KGSampler(dataset)`` reads ``dataset.head_entities / tail_entities / entity_num`` like the
    reference; arrays can be given directly instead.
    """

    def __init__(self, dataset=None, distribution="uniform", alpha=1.0, *, heads=None, tails=None, entity_num=None,
                 stream: MTStream | None = None, device="cuda"):
        if distribution != "uniform":
            raise NotImplementedError("only the uniform distribution is on the fused path")
        if dataset is not None:
            heads, tails, entity_num = dataset.head_entities, dataset.tail_entities, dataset.entity_num
        self.entity_num = int(entity_num)
        stream = stream if stream is not None else MTStream(device=device)
        super().__init__(heads, tails, self.entity_num, self.entity_num, stream, "head entities")

    def sample_by_entity_ids(self, head_entity_ids, num: int = 1) -> torch.Tensor:
        """[len(heads) * num] int64 on the device, j-major: out[j*len + i] (sampler.py:338-357)."""
        return self.sample_by_key_ids(head_entity_ids, num)


class RecSampler(_FilteredUniformSampler):
    """Item negatives filtered by the user's interacted items (one phase of sampler.py:186-291)."""

    def __init__(self, users, items, n_users: int, n_items: int, stream: MTStream | None = None, device="cuda"):
        stream = stream if stream is not None else MTStream(device=device)
        self.user_num, self.item_num = int(n_users), int(n_items)
        super().__init__(users, items, self.user_num, self.item_num, stream, "users")

    def sample_by_user_ids(self, user_ids, item_ids=None, num: int = 1) -> torch.Tensor:
        return self.sample_by_key_ids(user_ids, num)
