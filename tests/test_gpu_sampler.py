"""The GPU sampler (through the C ABI) against the reference's KGSampler / Sampler outputs
(tests/golden/sampler.npz), NumPy itself, and the oracle on larger inputs.  Bar: bit-exact ids
and bit-exact advanced MT19937 state."""

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import mt19937 as omt

pytestmark = pytest.mark.gpu


def _state_equal(stream, key, pos):
    st = stream.get_state()
    np.testing.assert_array_equal(st[1], key)
    assert st[2] == int(pos)


def test_seed_matches_numpy():
    from hopwise_b200.sampler import MTStream

    for seed in (0, 1, 2024, 2**32 - 1):
        np.random.seed(seed)
        st = np.random.get_state()
        _state_equal(MTStream(seed=seed), st[1], st[2])


def test_golden_reference_stream():
    from hopwise_b200.sampler import KGSampler, MTStream, RecSampler

    g = load_golden("sampler.npz")
    E, U, I = int(g["E"]), int(g["U"]), int(g["I"])
    stream = MTStream(state=("MT19937", g["state0_key"], int(g["state0_pos"])))
    kg = KGSampler(heads=g["heads"], tails=g["tails"], entity_num=E, stream=stream)
    rec = RecSampler(g["rec_users"], g["rec_items"], U, I, stream=stream)
    for c in range(int(g["n_calls"])):
        num = int(g[f"call{c}/num"])
        neg_t = kg.sample_by_entity_ids(g[f"call{c}/heads"], num)
        np.testing.assert_array_equal(neg_t.cpu().numpy(), g[f"call{c}/neg_tails"])
        _state_equal(stream, g[f"call{c}/kg_key"], g[f"call{c}/kg_pos"])
        neg_i = rec.sample_by_user_ids(g[f"call{c}/users"], None, num)
        np.testing.assert_array_equal(neg_i.cpu().numpy(), g[f"call{c}/neg_items"])
        _state_equal(stream, g[f"call{c}/rec_key"], g[f"call{c}/rec_pos"])
    assert not stream.exhausted()


@pytest.mark.parametrize("seed,E,n,num", [(2024, 34629, 2048, 1), (7, 300, 2048, 4), (3, 30001, 5000, 3), (11, 70, 33, 64)])
def test_against_numpy_and_oracle(seed, E, n, num):
    """NumPy's own generator drives an emulation of the reference loop (oracle), the device
    stream must produce the same ids and end in np.random's state."""
    from hopwise_b200.sampler import KGSampler, MTStream

    rng = np.random.default_rng(seed)
    n_tri = 20 * E if E < 1000 else 3 * E
    heads = rng.integers(1, E, n_tri)
    tails = rng.integers(1, E, n_tri)
    q = heads[rng.integers(0, n_tri, n)]
    off, vals = omt.build_used_csr(heads, tails, E)
    gen = omt.MT19937(seed)
    want = omt.sample_by_key_ids(gen, q, num, off, vals, 1, E)
    stream = MTStream(seed=seed)
    kg = KGSampler(heads=heads, tails=tails, entity_num=E, stream=stream)
    got = kg.sample_by_entity_ids(q, num)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    _state_equal(stream, gen.key, gen.pos)
    # a second call continues the stream
    want2 = omt.sample_by_key_ids(gen, q[:100], 1, off, vals, 1, E)
    np.testing.assert_array_equal(kg.sample_by_entity_ids(q[:100], 1).cpu().numpy(), want2)
    _state_equal(stream, gen.key, gen.pos)
    # and NumPy on the host can take over from the device state
    np.random.set_state(stream.get_state())
    a = np.random.randint(1, E, 10)
    np.testing.assert_array_equal(a, gen.randint(1, E, 10))


def test_unfiltered_draws_equal_numpy_randint():
    from hopwise_b200.sampler import KGSampler, MTStream

    E = 34629
    np.random.seed(5)
    want = np.random.randint(1, E, 5000)
    stream = MTStream(seed=5)
    kg = KGSampler(heads=np.array([1]), tails=np.array([0]), entity_num=E, stream=stream)  # nothing forbidden in range
    got = kg.sample_by_entity_ids(np.full(5000, 2), 1)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    st = np.random.get_state()
    _state_equal(stream, st[1], st[2])


def test_negatives_respect_the_filter_at_full_size():
    """BASELINE config 2 shape: 1M triples over 30k entities, 2048 x 64 negatives."""
    from hopwise_b200.sampler import KGSampler, MTStream

    E = 30001
    rng = np.random.default_rng(2024)
    heads = rng.integers(1, E, 1_000_000)
    tails = rng.integers(1, E, 1_000_000)
    kg = KGSampler(heads=heads, tails=tails, entity_num=E, stream=MTStream(seed=2024))
    q = heads[:2048]
    out = kg.sample_by_entity_ids(q, 64).cpu().numpy()
    assert out.shape == (2048 * 64,) and out.min() >= 1 and out.max() < E
    pairs = set(zip(heads.tolist(), tails.tolist()))
    keys = np.tile(q, 64)
    assert not any((h, t) in pairs for h, t in zip(keys[:20000].tolist(), out[:20000].tolist()))
    # identical to the oracle on the first call at this size
    off, vals = omt.build_used_csr(heads, tails, E)
    want = omt.sample_by_key_ids(omt.MT19937(2024), q, 64, off, vals, 1, E)
    np.testing.assert_array_equal(out, want)


def test_saturated_key_is_refused():
    from hopwise_b200.sampler import KGSampler

    E = 12
    heads = np.concatenate([np.full(E - 1, 3), [4]])
    tails = np.concatenate([np.arange(1, E), [5]])
    with pytest.raises(ValueError):
        KGSampler(heads=heads, tails=tails, entity_num=E)


# ---- popularity-biased candidates (sampler.py:68-116) ---------------------------------------------------------

@pytest.mark.parametrize("tag,alpha", [("a1", 1.0), ("a05", 0.5)])
def test_popularity_golden_reference_stream(tag, alpha):
    """KG and rec samplers in popularity mode on one stream: the reference's ids and MT19937 state, call by call."""
    from hopwise_b200.sampler import KGSampler, MTStream, RecSampler

    g = load_golden("sampler_pop.npz")
    E, U, I = int(g["E"]), int(g["U"]), int(g["I"])
    stream = MTStream(state=("MT19937", g[f"{tag}/state0_key"], int(g[f"{tag}/state0_pos"])))
    kg = KGSampler(heads=g["heads"], tails=g["tails"], entity_num=E, stream=stream, distribution="popularity", alpha=alpha)
    rec = RecSampler(g["rec_users"], g["rec_items"], U, I, stream=stream, distribution="popularity", alpha=alpha)
    np.testing.assert_array_equal(kg.pop[0].cpu().numpy(), g[f"{tag}/kg_keys"])
    np.testing.assert_array_equal(kg.pop[1].cpu().numpy(), g[f"{tag}/kg_prob"])
    for c in range(int(g[f"{tag}/n_calls"])):
        p = f"{tag}/call{c}/"
        num = int(g[p + "num"])
        neg_t = kg.sample_by_entity_ids(g[p + "heads"], num)
        np.testing.assert_array_equal(neg_t.cpu().numpy(), g[p + "neg_tails"])
        _state_equal(stream, g[p + "kg_key"], g[p + "kg_pos"])
        neg_i = rec.sample_by_user_ids(g[p + "users"], None, num)
        np.testing.assert_array_equal(neg_i.cpu().numpy(), g[p + "neg_items"])
        _state_equal(stream, g[p + "rec_key"], g[p + "rec_pos"])
    assert not stream.exhausted()


@pytest.mark.parametrize("seed,E,n,num,alpha", [(2024, 34629, 2048, 1, 1.0), (7, 300, 2048, 4, 0.5), (3, 30001, 5000, 3, 0.75)])
def test_popularity_against_oracle(seed, E, n, num, alpha):
    from hopwise_b200.sampler import KGSampler, MTStream

    rng = np.random.default_rng(seed)
    n_tri = 20 * E if E < 1000 else 3 * E
    heads = 1 + (rng.random(n_tri) ** 2 * (E - 1)).astype(np.int64)
    tails = 1 + (rng.random(n_tri) ** 3 * (E - 1)).astype(np.int64)
    q = heads[rng.integers(0, n_tri, n)]
    off, vals = omt.build_used_csr(heads, tails, E)
    table = omt.build_alias_table(np.concatenate([heads, tails]), alpha)
    gen = omt.MT19937(seed)
    want = omt.sample_by_key_ids(gen, q, num, off, vals, 1, E, pop_table=table)
    stream = MTStream(seed=seed)
    kg = KGSampler(heads=heads, tails=tails, entity_num=E, stream=stream, distribution="popularity", alpha=alpha)
    got = kg.sample_by_entity_ids(q, num)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    _state_equal(stream, gen.key, gen.pos)
    # the filter holds and NumPy can take over from the device state
    got = got.cpu().numpy()
    for j in (0, num - 1):
        for i in range(0, n, 97):
            assert got[j * n + i] not in set(vals[off[q[i]]:off[q[i] + 1]].tolist())
    np.random.set_state(stream.get_state())
    np.testing.assert_array_equal(np.random.random(5), omt.random_sample(gen, 5))


def test_popularity_single_key_table():
    """One candidate key: randint(0, 1, L) is constant and draws nothing, only the doubles advance the stream."""
    from hopwise_b200.sampler import KGSampler, MTStream

    heads, tails = np.array([3, 3, 3]), np.array([3, 3, 3])   # key 3 only; head 5 has no forbidden tails
    stream = MTStream(seed=4)
    kg = KGSampler(heads=heads, tails=tails, entity_num=9, stream=stream, distribution="popularity")
    out = kg.sample_by_entity_ids(np.array([5, 5, 6, 7]), 2)
    np.testing.assert_array_equal(out.cpu().numpy(), np.full(8, 3))
    np.random.seed(4)
    np.random.random(8)
    _state_equal(stream, np.random.get_state()[1], np.random.get_state()[2])
