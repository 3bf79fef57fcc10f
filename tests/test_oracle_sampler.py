"""Pin the MT19937 / masked-rejection / filtered-sampling oracle to numpy and to the
reference's KGSampler + Sampler outputs (tests/golden/sampler.npz)."""

import numpy as np
import pytest

from oracle.mt19937 import MT19937, build_used_csr, sample_by_key_ids

from conftest import load_golden


@pytest.mark.parametrize(
    "seed,low,high,n", [(2024, 1, 34629, 5000), (7, 1, 300, 3000), (1, 1, 65537, 2000), (5, 1, 2, 10), (9, 0, 1 << 20, 999)]
)
def test_randint_bit_exact_with_numpy(seed, low, high, n):
    np.random.seed(seed)
    ref_a = np.random.randint(low, high, n)
    ref_b = np.random.randint(low, high, 77)
    gen = MT19937(seed)
    np.testing.assert_array_equal(gen.randint(low, high, n), ref_a)
    np.testing.assert_array_equal(gen.randint(low, high, 77), ref_b)
    st = np.random.get_state()
    np.testing.assert_array_equal(st[1], gen.key)
    assert st[2] == gen.pos


def test_raw_words_match_numpy():
    np.random.seed(123)
    ref = np.random.randint(0, 1 << 32, 2000, dtype=np.uint64)  # rng == 0xFFFFFFFF: raw words
    gen = MT19937(123)
    np.testing.assert_array_equal(gen.next_words(2000).astype(np.uint64), ref)


def test_filtered_sampling_matches_reference_stream():
    g = load_golden("sampler.npz")
    E, U, I = int(g["E"]), int(g["U"]), int(g["I"])
    kg_off, kg_vals = build_used_csr(g["heads"], g["tails"], E)
    rec_off, rec_vals = build_used_csr(g["rec_users"], g["rec_items"], U)
    gen = MT19937()
    gen.set_state(("MT19937", g["state0_key"], int(g["state0_pos"])))
    for c in range(int(g["n_calls"])):
        num = int(g[f"call{c}/num"])
        neg_t = sample_by_key_ids(gen, g[f"call{c}/heads"], num, kg_off, kg_vals, 1, E)
        np.testing.assert_array_equal(neg_t, g[f"call{c}/neg_tails"])
        np.testing.assert_array_equal(gen.key, g[f"call{c}/kg_key"])
        assert gen.pos == int(g[f"call{c}/kg_pos"])
        neg_i = sample_by_key_ids(gen, g[f"call{c}/users"], num, rec_off, rec_vals, 1, I)
        np.testing.assert_array_equal(neg_i, g[f"call{c}/neg_items"])
        np.testing.assert_array_equal(gen.key, g[f"call{c}/rec_key"])
        assert gen.pos == int(g[f"call{c}/rec_pos"])


def test_negatives_never_in_used_set():
    g = load_golden("sampler.npz")
    E = int(g["E"])
    off, vals = build_used_csr(g["heads"], g["tails"], E)
    gen = MT19937(11)
    keys = g["heads"][:200]
    out = sample_by_key_ids(gen, keys, 4, off, vals, 1, E)
    for j in range(4):
        for i, k in enumerate(keys):
            assert out[j * len(keys) + i] not in set(vals[off[k] : off[k + 1]])
            assert 1 <= out[j * len(keys) + i] < E


def test_empty_request():
    gen = MT19937(3)
    before = gen.get_state()
    out = sample_by_key_ids(gen, np.zeros(0, dtype=np.int64), 1, np.zeros(2, np.int64), np.zeros(0, np.int64), 1, 5)
    assert out.shape == (0,)
    assert gen.pos == before[2]


# ---- popularity-biased candidates (sampler.py:68-116) ---------------------------------------------------------

def test_random_sample_bit_exact_with_numpy():
    from oracle.mt19937 import random_sample

    for seed, pre in ((5, 0), (6, 1), (2024, 623)):    # odd offsets: word pairs straddle the 624-word blocks
        np.random.seed(seed)
        np.random.randint(0, 1 << 30, pre)
        want = np.random.random(1500)
        gen = MT19937(seed)
        gen.randint(0, 1 << 30, pre)
        np.testing.assert_array_equal(random_sample(gen, 1500), want)
        st = np.random.get_state()
        np.testing.assert_array_equal(st[1], gen.key)
        assert st[2] == gen.pos


@pytest.mark.parametrize("tag,alpha", [("a1", 1.0), ("a05", 0.5)])
def test_alias_table_matches_reference(tag, alpha):
    """The oracle's dict restatement and the product's array builder (host set-up code, no GPU involved) against the
    table the reference built: same key order, bit-equal probabilities, same aliases."""
    from hopwise_b200.sampler import build_alias_table as product_table
    from oracle.mt19937 import build_alias_table

    g = load_golden("sampler_pop.npz")
    for name, cand in (("kg", np.concatenate([g["heads"], g["tails"]])), ("rec", g["rec_items"])):
        keys, prob, alias = build_alias_table(cand, alpha)
        np.testing.assert_array_equal(np.array(keys), g[f"{tag}/{name}_keys"])
        np.testing.assert_array_equal(np.array([prob[k] for k in keys]), g[f"{tag}/{name}_prob"])
        np.testing.assert_array_equal(np.array([alias[k] for k in keys]), g[f"{tag}/{name}_alias"])
        pk, pp, pa = product_table(cand, alpha)
        np.testing.assert_array_equal(pk, g[f"{tag}/{name}_keys"])
        np.testing.assert_array_equal(pp, g[f"{tag}/{name}_prob"])
        np.testing.assert_array_equal(pa, g[f"{tag}/{name}_alias"])


@pytest.mark.parametrize("tag,alpha", [("a1", 1.0), ("a05", 0.5)])
def test_popularity_sampling_matches_reference_stream(tag, alpha):
    from oracle.mt19937 import build_alias_table

    g = load_golden("sampler_pop.npz")
    E, U, I = int(g["E"]), int(g["U"]), int(g["I"])
    kg_off, kg_vals = build_used_csr(g["heads"], g["tails"], E)
    rec_off, rec_vals = build_used_csr(g["rec_users"], g["rec_items"], U)
    kg_tab = build_alias_table(np.concatenate([g["heads"], g["tails"]]), alpha)
    rec_tab = build_alias_table(g["rec_items"], alpha)
    gen = MT19937()
    gen.set_state(("MT19937", g[f"{tag}/state0_key"], int(g[f"{tag}/state0_pos"])))
    for c in range(int(g[f"{tag}/n_calls"])):
        p = f"{tag}/call{c}/"
        num = int(g[p + "num"])
        neg_t = sample_by_key_ids(gen, g[p + "heads"], num, kg_off, kg_vals, 1, E, pop_table=kg_tab)
        np.testing.assert_array_equal(neg_t, g[p + "neg_tails"])
        np.testing.assert_array_equal(gen.key, g[p + "kg_key"])
        assert gen.pos == int(g[p + "kg_pos"])
        neg_i = sample_by_key_ids(gen, g[p + "users"], num, rec_off, rec_vals, 1, I, pop_table=rec_tab)
        np.testing.assert_array_equal(neg_i, g[p + "neg_items"])
        np.testing.assert_array_equal(gen.key, g[p + "rec_key"])
        assert gen.pos == int(g[p + "rec_pos"])


def test_alias_table_large_and_single_key():
    """The array builder against the dict restatement on a long-tailed table, and the one-key table (randint(0, 1)
    consumes no words)."""
    from hopwise_b200.sampler import build_alias_table as product_table
    from oracle.mt19937 import build_alias_table, pop_sampling

    rng = np.random.default_rng(3)
    cand = 1 + (rng.random(40000) ** 4 * 4999).astype(np.int64)
    for alpha in (1.0, 0.75, 0.0):
        keys, prob, alias = build_alias_table(cand, alpha)
        pk, pp, pa = product_table(cand, alpha)
        np.testing.assert_array_equal(pk, np.array(keys))
        np.testing.assert_array_equal(pp, np.array([prob[k] for k in keys]))
        np.testing.assert_array_equal(pa, np.array([alias[k] for k in keys]))
    one = build_alias_table([7, 7, 7], 1.0)
    gen = MT19937(1)
    np.random.seed(1)
    np.random.random(4)   # the only words drawn are the doubles'
    np.testing.assert_array_equal(pop_sampling(gen, one, 4), [7, 7, 7, 7])
    assert np.random.get_state()[2] == gen.pos
