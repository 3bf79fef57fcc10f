"""Randomised cross-check of the oracle against the LIVE reference (hopwise imported from /root/reference).

The golden fixtures pin the oracle on two fixed shapes; this fuzzes it on many: random table sizes (odd and
even embedding sizes), ragged rec / KG halves, duplicate rows, K > 1 tiled negatives, and random KGs for the
sampler.  Runs only where the reference tree exists (the build container); it never runs on the GPU box.
"""

import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from _ref_harness import REF_CONFIG, FakeDataset, import_reference, reference_available  # noqa: E402

from kge_helpers import make_oracle_model, random_batch, tile_batch, to_cpu_batch  # noqa: E402
from oracle import mt19937 as omt  # noqa: E402

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")

MODELS = ("TransE", "RotatE", "DistMult", "ComplEx", "TorusE", "TransH", "TransD")


def _ref_classes():
    import_reference()
    from hopwise.data.interaction import Interaction
    from hopwise.model.knowledge_graph_embedding_recommender.complex import ComplEx
    from hopwise.model.knowledge_graph_embedding_recommender.distmult import DistMult
    from hopwise.model.knowledge_graph_embedding_recommender.rotate import RotatE
    from hopwise.model.knowledge_graph_embedding_recommender.toruse import TorusE
    from hopwise.model.knowledge_graph_embedding_recommender.transd import TransD
    from hopwise.model.knowledge_graph_embedding_recommender.transe import TransE
    from hopwise.model.knowledge_graph_embedding_recommender.transh import TransH

    return {"TransE": TransE, "RotatE": RotatE, "DistMult": DistMult, "ComplEx": ComplEx, "TorusE": TorusE,
            "TransH": TransH, "TransD": TransD}, Interaction


@pytest.mark.parametrize("name", MODELS)
@pytest.mark.parametrize("case", range(4))
def test_loss_gradients_and_scores_match_the_reference(name, case):
    classes, Interaction = _ref_classes()
    rng = np.random.default_rng(100 * case + len(name))
    U, I, E, R = (int(x) for x in (rng.integers(5, 40), rng.integers(4, 30), 0, rng.integers(4, 9)))
    E = I + int(rng.integers(1, 50))
    d = int(rng.choice([3, 8, 12, 17, 33]))
    k = int(rng.choice([1, 1, 3]))
    margin = float(rng.choice([0.5, 1.0, 2.0]))
    torch.manual_seed(7 + case)
    ref = classes[name](dict(REF_CONFIG, embedding_size=d, margin=margin), FakeDataset(U, I, E, R))
    ora = make_oracle_model(name, U, I, E, R, d, margin=margin)
    ora.load_state_dict(ref.state_dict(), strict=True)   # same key names, same tensors

    b = tile_batch(random_batch(rng, U, I, E, R, int(rng.integers(1, 20)), int(rng.integers(1, 20)), k, k), k, k)
    inter = Interaction({key: torch.as_tensor(v, dtype=torch.long) for key, v in b.items()})
    loss_r = ref.calculate_loss(inter)
    loss_o = ora.calculate_loss(to_cpu_batch(b))
    np.testing.assert_allclose(loss_o.item(), loss_r.item(), rtol=1e-6)
    loss_r.backward()
    loss_o.backward()
    for (kr, pr), (ko, po) in zip(ref.named_parameters(), ora.named_parameters()):
        assert kr == ko
        gr = pr.grad.numpy() if pr.grad is not None else np.zeros_like(pr.detach().numpy())
        go = po.grad.numpy() if po.grad is not None else np.zeros_like(gr)
        # (TransH: the reference looks the hyperplane row up twice per projection, so its autograd adds the two
        # paths' fp32 contributions in another order; elements that nearly cancel differ by ~1e-8 absolute)
        np.testing.assert_allclose(go, gr, rtol=1e-5, atol=5e-8 if name in ("TransH", "TransD") else 1e-8, err_msg=kr)

    with torch.no_grad():
        np.testing.assert_allclose(ora.predict(to_cpu_batch(b)).numpy(), ref.predict(inter).numpy(), rtol=1e-5, atol=1e-6)
        has_kg = hasattr(ref, "predict_kg")   # transh.py scores users against items only
        if has_kg:
            np.testing.assert_allclose(ora.predict_kg(to_cpu_batch(b)).numpy(), ref.predict_kg(inter).numpy(), rtol=1e-5,
                                       atol=1e-6)
        users = torch.as_tensor(rng.integers(1, U, 6), dtype=torch.long)
        fr = ref.full_sort_predict(Interaction({"user_id": users})).view(-1, I).numpy()
        fo = ora.full_sort_predict({"user_id": users}).numpy()
        np.testing.assert_allclose(fo, fr, rtol=1e-5, atol=1e-6)
        if has_kg and name != "TransD":   # (transd.py:192-217 projects the head with <h, h>: not mirrored)
            kb = {"head_id": torch.as_tensor(b["head_id"][:5], dtype=torch.long),
                  "relation_id": torch.as_tensor(b["relation_id"][:5], dtype=torch.long)}
            fr = ref.full_sort_predict_kg(Interaction(kb)).view(-1, E).numpy()
            fo = ora.full_sort_predict_kg(kb).numpy()
            np.testing.assert_allclose(fo, fr, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", range(5))
def test_kg_sampler_stream_matches_the_reference(case):
    import_reference()
    from hopwise.sampler import KGSampler

    rng = np.random.default_rng(500 + case)
    E = int(rng.integers(8, 300))
    n_tri = int(rng.integers(20, 2000))
    heads, tails = rng.integers(1, E, n_tri), rng.integers(1, E, n_tri)
    if case % 2:   # a hub head linked to most entities: several rejection rounds
        hub = rng.permutation(np.arange(1, E))[: max(1, E - 4)]
        heads, tails = np.concatenate([heads, np.full(len(hub), 3)]), np.concatenate([tails, hub])
    kg = KGSampler(FakeDataset(10, 5, E, 6, heads=heads, tails=tails))
    off, vals = omt.build_used_csr(heads, tails, E)
    seed = 2024 + case
    np.random.seed(seed)
    gen = omt.MT19937(seed)
    for num in (1, 4, 1):   # consecutive calls share the stream, as in a training run
        q = heads[rng.integers(0, len(heads), int(rng.integers(1, 200)))]
        want = kg.sample_by_entity_ids(q, num).numpy()
        got = omt.sample_by_key_ids(gen, q, num, off, vals, 1, E)
        np.testing.assert_array_equal(got, want)
    # and the generator the reference leaves behind continues like the oracle's (numpy refills its block
    # lazily, so the next draws are compared rather than the raw state words)
    np.testing.assert_array_equal(gen.randint(0, 1 << 30, 700), np.random.randint(0, 1 << 30, 700))


@pytest.mark.parametrize("case", range(4))
def test_collector_and_metrics_match_the_reference(case):
    """Masking + top-k + hit matrix + Recall/MRR/NDCG/Hit/Precision of oracle/fullsort.py against the reference's
    Collector / Evaluator / metric classes on random scores (continuous scores: no ties to disagree on)."""
    import_reference()
    from hopwise.evaluator import Collector, Evaluator
    from hopwise.evaluator.register import metrics_dict

    from oracle import fullsort as fs

    rng = np.random.default_rng(900 + case)
    n, I = int(rng.integers(3, 40)), int(rng.integers(30, 200))
    k = int(rng.integers(2, 21))
    scores = rng.standard_normal((n, I)).astype(np.float32)
    hist_u, hist_i, pos_u, pos_i = [], [], [], []
    for u in range(n):
        perm = rng.permutation(np.arange(1, I))
        nh, npos = int(rng.integers(0, I // 3)), int(rng.integers(1, 6))
        hist_u += [u] * nh
        hist_i += list(perm[:nh])
        pos_u += [u] * npos
        pos_i += list(perm[nh:nh + npos])
    cfg = {"eval_args": {"mode": "full"}, "topk": [max(1, k // 2), k], "device": torch.device("cpu"),
           "metrics": ["Recall", "MRR", "NDCG", "Hit", "Precision"], "metric_decimal_place": 4, "tsne": None}
    coll = Collector(cfg)
    s = torch.from_numpy(scores.copy())
    s[:, 0] = -np.inf
    if hist_u:
        s[(torch.as_tensor(hist_u), torch.as_tensor(hist_i))] = -np.inf
    coll.eval_batch_collect(s, None, torch.as_tensor(pos_u), torch.as_tensor(pos_i))
    struct = coll.get_data_struct()
    want_rec = struct.get("rec.topk").numpy()
    want = Evaluator(cfg).evaluate(struct)

    masked = fs.mask_scores(scores, np.array(hist_u, dtype=np.int64), np.array(hist_i, dtype=np.int64))
    ids, _ = fs.topk_canonical(masked, k)
    np.testing.assert_array_equal(ids, torch.topk(s, k, dim=-1)[1].numpy())
    rec = fs.hits(ids, np.array(pos_u), np.array(pos_i), I)
    np.testing.assert_array_equal(rec, want_rec)
    got = fs.metric_values(rec, topk=tuple(cfg["topk"]), decimals=4)
    for name, v in want.items():
        assert got[name] == v, name
    mats = fs.metric_matrices(rec[:, :k], rec[:, k])
    for name in ("recall", "mrr", "ndcg", "hit", "precision"):
        m = metrics_dict[name](cfg)
        ref = m.metric_info(rec[:, :k].astype(bool), rec[:, k]) if name in ("recall", "ndcg") else m.metric_info(
            rec[:, :k].astype(bool))
        np.testing.assert_array_equal(mats[name], np.asarray(ref, dtype=np.float64), err_msg=name)


class _RecDataset:
    """What hopwise's rec-side Sampler reads from a dataset (sampler.py:199-252)."""

    def __init__(self, users, items, n_users, n_items):
        self.uid_field, self.iid_field = "user_id", "item_id"
        self.user_num, self.item_num = n_users, n_items
        self.inter_feat = {"user_id": torch.as_tensor(users), "item_id": torch.as_tensor(items)}


@pytest.mark.parametrize("case", range(4))
def test_rec_and_kg_samplers_share_one_stream_like_the_reference(case):
    """One training step draws KG negatives first, then rec negatives, from numpy's one global generator
    (knowledge_dataloader.py:137-145); the oracle reproduces both with one MT19937 stream."""
    import_reference()
    from hopwise.sampler import KGSampler, Sampler

    rng = np.random.default_rng(700 + case)
    U, I = int(rng.integers(5, 60)), int(rng.integers(6, 80))
    E = I + int(rng.integers(1, 100))
    heads, tails = rng.integers(1, E, 600), rng.integers(1, E, 600)
    ru, ri = rng.integers(1, U, 400), rng.integers(1, I, 400)
    # no user may own every item (the reference raises then)
    keep = np.ones(len(ru), dtype=bool)
    for u in range(1, U):
        if len(set(ri[ru == u])) >= I - 2:
            keep &= ru != u
    ru, ri = ru[keep], ri[keep]
    kg = KGSampler(FakeDataset(U, I, E, 6, heads=heads, tails=tails))
    rec = Sampler("train", _RecDataset(ru, ri, U, I)).set_phase("train")
    kg_off, kg_vals = omt.build_used_csr(heads, tails, E)
    rec_off, rec_vals = omt.build_used_csr(ru, ri, U)
    seed = 31 + case
    np.random.seed(seed)
    gen = omt.MT19937(seed)
    for num in (1, 2, 1):
        qh = heads[rng.integers(0, len(heads), int(rng.integers(1, 120)))]
        qu = ru[rng.integers(0, len(ru), int(rng.integers(1, 120)))]
        np.testing.assert_array_equal(omt.sample_by_key_ids(gen, qh, num, kg_off, kg_vals, 1, E),
                                      kg.sample_by_entity_ids(qh, num).numpy())
        np.testing.assert_array_equal(omt.sample_by_key_ids(gen, qu, num, rec_off, rec_vals, 1, I),
                                      rec.sample_by_user_ids(qu, None, num).numpy())
