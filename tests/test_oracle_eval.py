"""Pin the full-sort / top-k / metric oracle to the reference's Collector + Evaluator output
and to the reference's own known-answer vectors
(/root/reference/tests/metrics/test_topk_metrics.py:26-108, restated below)."""

import numpy as np

from oracle import fullsort as fs

from conftest import load_golden

# the reference test's inputs: a 4x3 hit matrix and the positives per user
POS_IDX = np.array([[0, 0, 0], [1, 1, 1], [1, 0, 1], [0, 0, 1]])
POS_LEN = np.array([1, 3, 4, 2])


def test_reference_known_answers():
    m = fs.metric_matrices(POS_IDX, POS_LEN)
    assert m["hit"].tolist() == [[0, 0, 0], [1, 1, 1], [1, 1, 1], [0, 0, 1]]
    assert m["mrr"].tolist() == [[0, 0, 0], [1, 1, 1], [1, 1, 1], [0, 0, 1 / 3]]
    assert m["recall"].tolist() == [[0, 0, 0], [1 / 3, 2 / 3, 3 / 3], [1 / 4, 1 / 4, 2 / 4], [0, 0, 1 / 2]]
    assert m["precision"].tolist() == [[0, 0, 0], [1 / 1, 2 / 2, 3 / 3], [1 / 1, 1 / 2, 2 / 3], [0, 0, 1 / 3]]
    l2 = np.log2
    want = [
        [0, 0, 0],
        [1, 1, 1],
        [1, (1 / l2(2) / (1 / l2(2) + 1 / l2(3))), ((1 / l2(2) + 1 / l2(4)) / (1 / l2(2) + 1 / l2(3) + 1 / l2(4)))],
        [0, 0, (1 / l2(4) / (1 / l2(2) + 1 / l2(3)))],
    ]
    assert m["ndcg"].tolist() == np.array(want).tolist()


def test_collector_and_evaluator_golden():
    g = load_golden("eval.npz")
    k = int(g["k"])
    masked = fs.mask_scores(g["scores"], g["hist_u"], g["hist_i"])
    ids, _ = fs.topk_canonical(masked, k)
    # user 3 has fewer than k unmasked items: the tail of its list is -inf ties, whose order
    # torch leaves unspecified; compare it only on the finite prefix
    finite = np.isfinite(np.take_along_axis(masked, ids, axis=1))
    ref_ids = g["topk_ids"]
    assert (ids[finite] == ref_ids[finite]).all()
    assert not finite[3].all() and finite[np.arange(len(finite)) != 3].all()
    rec_topk = fs.hits(ids, g["pos_u"], g["pos_i"], masked.shape[1])
    np.testing.assert_array_equal(rec_topk, g["rec_topk"])
    mats = fs.metric_matrices(rec_topk[:, :k], rec_topk[:, k])
    for name in ("recall", "mrr", "ndcg", "hit", "precision"):
        np.testing.assert_array_equal(mats[name], g["matrix/" + name])
    vals = fs.metric_values(rec_topk, topk=(5, k), decimals=4)
    for name, v in zip(g["metric_names"], g["metric_values"]):
        assert vals[str(name)] == v


def test_canonical_tie_order():
    s = np.array([[1, 3, 3, 3, 2, 3, -np.inf, 3]], dtype=np.float32)
    ids, _ = fs.topk_canonical(s, 3)
    assert ids.tolist() == [[1, 2, 3]]
