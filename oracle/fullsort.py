"""Masked full-sort, canonical top-k, positive-hit matrix and ranking metrics
(oracle; test infrastructure only).

What it follows (relative to /root/reference/hopwise/):
  trainer/trainer.py:716-735   scores.view(-1, I); scores[:, 0] = -inf; scores[history] = -inf
  evaluator/collector.py:176-183  topk(scores, max(topk)); pos_matrix[pos_u, pos_i] = 1;
                                  pos_len = row sums; pos_idx = gather(pos_matrix, topk_idx);
                                  rec.topk = cat(pos_idx, pos_len)  (int32 [n, k+1])
  evaluator/base_metric.py:75-99  split into bool pos_index / pos_len; mean over users; value at k-1
  evaluator/metrics.py:67-69 (Hit), 93-101 (MRR), 164-165 (Recall), 191-207 (NDCG), 231-232 (Precision)

Tie order: torch.topk's order among equal scores is implementation-defined, so this project
fixes the canonical order (score descending, item id ascending) -- SURVEY.md H6.

The metric formulas are pinned by the reference's known-answer vectors
(tests/metrics/test_topk_metrics.py:26-108), reproduced in tests/test_oracle_metrics.py.
"""

from __future__ import annotations

import numpy as np


def mask_scores(scores: np.ndarray, hist_u=None, hist_i=None) -> np.ndarray:
    out = np.array(scores, copy=True)
    out[:, 0] = -np.inf
    if hist_u is not None and len(hist_u):
        out[np.asarray(hist_u), np.asarray(hist_i)] = -np.inf
    return out


def topk_canonical(scores: np.ndarray, k: int):
    """ids [n,k] ordered by (score desc, id asc), and the scores at those ids."""
    n, I = scores.shape
    ids = np.broadcast_to(np.arange(I), (n, I))
    # lexsort: last key is primary
    order = np.lexsort((ids, -scores), axis=1)[:, :k]
    return order, np.take_along_axis(scores, order, axis=1)


def hits(topk_ids: np.ndarray, pos_u, pos_i, n_items: int) -> np.ndarray:
    """int32 [n, k+1]: pos_idx columns then pos_len (collector.py:178-183)."""
    n, k = topk_ids.shape
    pos = np.zeros((n, n_items), dtype=np.int32)
    pos[np.asarray(pos_u), np.asarray(pos_i)] = 1
    pos_idx = np.take_along_axis(pos, topk_ids, axis=1)
    return np.concatenate([pos_idx, pos.sum(1, keepdims=True)], axis=1).astype(np.int32)


def metric_matrices(pos_index: np.ndarray, pos_len: np.ndarray) -> dict:
    """Per-user [n,k] float64 matrices for recall / ndcg / mrr / hit / precision."""
    pos_index = np.asarray(pos_index).astype(bool)
    pos_len = np.asarray(pos_len)
    n, k = pos_index.shape
    ranks = np.arange(1, k + 1, dtype=np.float64)
    cum = np.cumsum(pos_index, axis=1)
    out = {
        "hit": (cum > 0).astype(int),
        "recall": cum / pos_len.reshape(-1, 1),
        "precision": cum / ranks,
    }
    first = pos_index.argmax(axis=1)
    any_hit = pos_index[np.arange(n), first]
    cols = np.arange(k)[None, :]
    rr = np.where(any_hit, 1.0 / (first + 1), 0.0)
    out["mrr"] = np.where(cols >= first[:, None], rr[:, None], 0.0)
    disc = 1.0 / np.log2(ranks + 1)
    dcg = np.cumsum(np.where(pos_index, disc, 0.0), axis=1)
    idcg_all = np.cumsum(disc)
    ilen = np.minimum(pos_len, k)
    col_eff = np.minimum(cols, ilen[:, None] - 1)  # freeze idcg after min(pos_len,k) terms
    out["ndcg"] = dcg / idcg_all[col_eff]
    return out


def metric_values(rec_topk: np.ndarray, topk=(10,), decimals: int | None = None) -> dict:
    """{'recall@10': ...} from the [n, kmax+1] rec.topk matrix (base_metric.py:75-99)."""
    kmax = rec_topk.shape[1] - 1
    mats = metric_matrices(rec_topk[:, :kmax], rec_topk[:, kmax])
    res = {}
    for name in ("recall", "mrr", "ndcg", "hit", "precision"):
        avg = mats[name].mean(axis=0)
        for k in topk:
            v = avg[k - 1]
            res[f"{name}@{k}"] = round(v, decimals) if decimals is not None else v
    return res
