"""MT19937 + numpy's masked-rejection ``randint`` + hopwise's filtered negative
sampling loop, restated from the published algorithms (oracle; test infrastructure only).

What it follows:
  * hopwise/sampler/sampler.py:140-183  AbstractSampler.sample_by_key_ids (both branches
    consume the stream identically: draw len(check_list) values, keep re-drawing the
    positions whose value is in used_ids[key])
  * hopwise/sampler/sampler.py:315-316  KGSampler._uni_sampling = np.random.randint(1, entity_num, n)
  * hopwise/sampler/sampler.py:226-227  Sampler._uni_sampling   = np.random.randint(1, item_num, n)
  * hopwise/sampler/sampler.py:321-336  used_ids[h] = set of all tails of head h (relation ignored)
  * numpy (third party, pinned 2.1.3 in uv.lock:1707-1708; 2.3.5 here) legacy
    RandomState.randint -> _rand_int64 -> random_bounded_uint64_fill with masked rejection:
    for rng = high-1-low <= 0xFFFFFFFF and rng != 0xFFFFFFFF:
        mask = smallest (2^b - 1) >= rng;  repeat v = next_uint32() & mask until v <= rng;  out = low + v
  * MT19937 (Matsumoto & Nishimura 1998): 624-word state, twist with MATRIX_A=0x9908b0df,
    tempering (11, 7/0x9d2c5680, 15/0xefc60000, 18); init_genrand(s) with multiplier 1812433253.
    np.random.seed(int) (legacy seeding) is init_genrand(seed) with pos=624; array seeds use
    init_by_array.

`MT19937.get_state()/set_state()` speak numpy's ('MT19937', key[624], pos, 0, 0.0) tuples, so a
test can hand the state to ``np.random.set_state`` and back.
"""

from __future__ import annotations

import numpy as np

N, M = 624, 397
MATRIX_A = np.uint32(0x9908B0DF)
UPPER, LOWER = np.uint32(0x80000000), np.uint32(0x7FFFFFFF)


def _init_genrand(s: int) -> np.ndarray:
    mt = np.zeros(N, dtype=np.uint64)
    mt[0] = s & 0xFFFFFFFF
    for i in range(1, N):
        mt[i] = (1812433253 * (int(mt[i - 1]) ^ (int(mt[i - 1]) >> 30)) + i) & 0xFFFFFFFF
    return mt.astype(np.uint32)


def _init_by_array(key) -> np.ndarray:
    mt = [int(x) for x in _init_genrand(19650218)]
    i, j = 1, 0
    klen = len(key)
    for _ in range(max(N, klen)):
        mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525)) + int(key[j]) + j) & 0xFFFFFFFF
        i += 1
        j += 1
        if i >= N:
            mt[0] = mt[N - 1]
            i = 1
        if j >= klen:
            j = 0
    for _ in range(N - 1):
        mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941)) - i) & 0xFFFFFFFF
        i += 1
        if i >= N:
            mt[0] = mt[N - 1]
            i = 1
    mt[0] = 0x80000000
    return np.array(mt, dtype=np.uint32)


def twist(mt: np.ndarray) -> np.ndarray:
    """Next 624-word state.  Three data-parallel phases (227 / 227 / 170 words): words
    [0,227) need only old words, [227,454) need new [0,227), [454,624) need new [227,397)
    (and word 623 needs new word 0) -- the same split the CUDA kernel uses."""
    old = mt
    new = np.empty_like(old)

    def mix(u, v, far):
        y = (u & UPPER) | (v & LOWER)
        return far ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), MATRIX_A, np.uint32(0))

    new[0:227] = mix(old[0:227], old[1:228], old[397:624])
    new[227:454] = mix(old[227:454], old[228:455], new[0:227])
    new[454:623] = mix(old[454:623], old[455:624], new[227:396])
    new[623] = mix(old[623:624], new[0:1], new[396:397])[0]
    return new


def temper(y: np.ndarray) -> np.ndarray:
    y = y ^ (y >> np.uint32(11))
    y = y ^ ((y << np.uint32(7)) & np.uint32(0x9D2C5680))
    y = y ^ ((y << np.uint32(15)) & np.uint32(0xEFC60000))
    return y ^ (y >> np.uint32(18))


class MT19937:
    def __init__(self, seed: int | None = None):
        self.key = np.zeros(N, dtype=np.uint32)
        self.pos = N
        if seed is not None:
            self.seed(seed)

    def seed(self, seed: int):
        # np.random.seed(int) (legacy seeding) = init_genrand(seed), pos = 624
        seed = int(seed)
        if not 0 <= seed <= 0xFFFFFFFF:
            raise ValueError("legacy integer seed must fit 32 bits")
        self.key = _init_genrand(seed)
        self.pos = N

    def seed_by_array(self, words):
        # np.random.seed([w0, w1, ...]) = init_by_array
        self.key = _init_by_array(list(words))
        self.pos = N

    def get_state(self):
        return ("MT19937", self.key.copy(), int(self.pos), 0, 0.0)

    def set_state(self, state):
        self.key = np.array(state[1], dtype=np.uint32).copy()
        self.pos = int(state[2])

    def next_words(self, count: int) -> np.ndarray:
        """`count` tempered 32-bit outputs, advancing the stream."""
        out = np.empty(count, dtype=np.uint32)
        done = 0
        while done < count:
            if self.pos >= N:
                self.key = twist(self.key)
                self.pos = 0
            take = min(count - done, N - self.pos)
            out[done : done + take] = temper(self.key[self.pos : self.pos + take])
            self.pos += take
            done += take
        return out

    def randint(self, low: int, high: int, size: int) -> np.ndarray:
        """np.random.randint(low, high, size) (int64 output), masked rejection."""
        rng = high - 1 - low
        out = np.empty(size, dtype=np.int64)
        if size == 0:
            return out
        if rng == 0:
            out[:] = low
            return out
        if not (0 < rng < 0xFFFFFFFF):
            raise NotImplementedError("only the 32-bit masked path is on hopwise's hot path")
        mask = np.uint32((1 << int(rng).bit_length()) - 1)
        filled = 0
        while filled < size:
            # never read past the word that yields the last needed value: take at most the
            # rest of the current block, stop at the (size-filled)-th accepted word
            if self.pos >= N:
                self.key = twist(self.key)
                self.pos = 0
            block = temper(self.key[self.pos :]) & mask
            ok = block <= np.uint32(rng)
            need = size - filled
            csum = np.cumsum(ok)
            if csum[-1] >= need:
                last = int(np.searchsorted(csum, need))  # index of the need-th accepted word
                vals = block[: last + 1][ok[: last + 1]]
                self.pos += last + 1
            else:
                vals = block[ok]
                self.pos = N
            out[filled : filled + len(vals)] = vals.astype(np.int64) + low
            filled += len(vals)
        return out


def random_sample(gen: "MT19937", size: int) -> np.ndarray:
    """np.random.random(size) on the legacy stream: per double two words, (w0 >> 5) * 2^26 + (w1 >> 6) over 2^53
    (numpy/random/src/mt19937/mt19937.h mt19937_next_double; numpy is a third-party dependency of the reference,
    the call site is sampler.py:112)."""
    w = gen.next_words(2 * size).astype(np.uint64)
    a, b = w[0::2] >> np.uint64(5), w[1::2] >> np.uint64(6)
    return (a.astype(np.float64) * 67108864.0 + b.astype(np.float64)) / 9007199254740992.0


def build_alias_table(candidates, alpha: float):
    """sampler.py:68-100 restated with the reference's own containers (dict in first-occurrence order, Python
    floats, FIFO lists); returns (keys list, prob dict, alias dict)."""
    cand = [int(x) for x in candidates]
    prob = {}
    for c in cand:
        prob[c] = prob.get(c, 0) + 1
    alias = {}
    for i in prob:
        alias[i] = -1
        prob[i] = pow(prob[i] / len(cand), alpha)
    norm = sum(prob.values())
    large_q, small_q = [], []
    for i in prob:
        prob[i] = prob[i] / norm * len(prob)
        if prob[i] > 1:
            large_q.append(i)
        elif prob[i] < 1:
            small_q.append(i)
    while large_q and small_q:
        lq, sq = large_q.pop(0), small_q.pop(0)
        alias[sq] = lq
        prob[lq] = prob[lq] - (1 - prob[sq])
        if prob[lq] < 1:
            small_q.append(lq)
        elif prob[lq] > 1:
            large_q.append(lq)
    return list(prob.keys()), prob, alias


def pop_sampling(gen: "MT19937", table, size: int) -> np.ndarray:
    """sampler.py:102-116: randint(0, n_keys, size), then random(size), then the alias rule."""
    keys, prob, alias = table
    idx = gen.randint(0, len(keys), size)
    p = random_sample(gen, size)
    return np.array([keys[i] if prob[keys[i]] > q else alias[keys[i]] for i, q in zip(idx, p)], dtype=np.int64)


def build_used_csr(keys: np.ndarray, values: np.ndarray, n_keys: int):
    """CSR of sorted, de-duplicated values per key: the array form of the reference's
    `used_ids[key] = set(values)` (sampler.py:229-252, 321-336)."""
    keys = np.asarray(keys, dtype=np.int64)
    values = np.asarray(values, dtype=np.int64)
    pairs = np.unique(np.stack([keys, values], axis=1), axis=0) if len(keys) else np.zeros((0, 2), np.int64)
    counts = np.bincount(pairs[:, 0], minlength=n_keys)
    off = np.zeros(n_keys + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    return off, pairs[:, 1].copy()


def _combined(off, vals):
    """(key, value) pairs of the CSR packed into one sorted int64 array."""
    seg_key = np.repeat(np.arange(len(off) - 1, dtype=np.int64), np.diff(off))
    return (seg_key << 32) | vals


def _member(combined, keys, cand):
    """cand[i] in used[keys[i]] for every i (one binary search over the packed pairs)."""
    probe = (keys.astype(np.int64) << 32) | cand.astype(np.int64)
    j = np.searchsorted(combined, probe)
    hit = np.zeros(len(probe), dtype=bool)
    inb = j < len(combined)
    hit[inb] = combined[j[inb]] == probe[inb]
    return hit


def sample_by_key_ids(gen: MT19937, key_ids, num: int, off, vals, low: int, high: int, pop_table=None) -> np.ndarray:
    """Filtered negatives, j-major layout out[j*len(keys)+i] (sampler.py:140-183); candidates are uniform in
    [low, high) or, with `pop_table` = build_alias_table(...), popularity-biased (sampler.py:102-116)."""
    key_ids = np.asarray(key_ids, dtype=np.int64)
    keys = np.tile(key_ids, num)
    total = len(keys)
    out = np.zeros(total, dtype=np.int64)
    check = np.arange(total)
    combined = _combined(np.asarray(off), np.asarray(vals))
    while len(check) > 0:
        if pop_table is None:
            out[check] = gen.randint(low, high, len(check))
        else:
            out[check] = pop_sampling(gen, pop_table, len(check))
        check = check[_member(combined, keys[check], out[check])]
    return out
