"""Fused full-sort evaluation: scores -> mask -> top-k -> hit matrix -> ranking metrics.

Host-side mirror of (paths under /root/reference/hopwise/):
  trainer/trainer.py:716-735      _full_sort_batch_eval (scores[:,0] = -inf, history -> -inf)
  evaluator/collector.py:152-206  Collector.eval_batch_collect ("rec.topk" = pos_idx | pos_len)
  evaluator/collector.py:234-249  get_data_struct
  evaluator/base_metric.py:75-99  used_info / topk_result (mean over users, value at k-1, rounding)
  evaluator/metrics.py:44-232     Hit, MRR, Recall, NDCG, Precision
  evaluator/evaluator.py:27-41    Evaluator.evaluate -> OrderedDict

The dense [n_users, n_items] score matrix of the reference never exists here: the kernel keeps
the per-user top-k on chip.  Tie order is (score desc, item id asc).
"""

from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch

from . import _abi

METRIC_ORDER = ("recall", "mrr", "ndcg", "hit", "precision")  # row order of kge_topk_metric_sums


def csr_from_pairs(rows, cols, n_rows: int, device):
    """(offsets[n_rows+1], cols sorted per row) from COO pairs such as the loader's
    (history_row, history_col) / (positive_u, positive_i) (general_dataloader.py:253-267)."""
    rows = torch.as_tensor(rows, dtype=torch.int64).to(device)
    cols = torch.as_tensor(cols, dtype=torch.int64).to(device)
    if rows.numel():
        span = int(cols.max().item()) + 1
        packed, _ = torch.sort(rows * span + cols)
        r2 = torch.div(packed, span, rounding_mode="floor")
        c2 = packed - r2 * span
        counts = torch.bincount(r2, minlength=n_rows)
    else:
        c2 = torch.zeros(0, dtype=torch.int64, device=device)
        counts = torch.zeros(n_rows, dtype=torch.int64, device=device)
    off = torch.zeros(n_rows + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts, 0, out=off[1:])
    return off, c2.contiguous()


def topk_hits(ids: torch.Tensor, pos_off: torch.Tensor, pos_items: torch.Tensor) -> torch.Tensor:
    """int32 [n, k+1]: hit flags of the top-k ids then pos_len (collector.py:178-183)."""
    n, k = ids.shape
    out = torch.empty(n, k + 1, dtype=torch.int32, device=ids.device)
    _abi.check(
        _abi.lib().kge_topk_hits(ids.data_ptr(), n, k, pos_off.data_ptr(), pos_items.data_ptr(), out.data_ptr(),
                                 _abi.stream_ptr()),
        "kge_topk_hits",
    )
    return out


def topk_metric_sums(rec_topk: torch.Tensor) -> torch.Tensor:
    """float64 [5, k] sums over users of recall, mrr, ndcg, hit, precision at every cutoff."""
    n, k1 = rec_topk.shape
    k = k1 - 1
    sums = torch.zeros(5, k, dtype=torch.float64, device=rec_topk.device)
    rec_topk = rec_topk.contiguous()
    _abi.check(
        _abi.lib().kge_topk_metric_sums(rec_topk.data_ptr(), n, k, sums.data_ptr(), _abi.stream_ptr()),
        "kge_topk_metric_sums",
    )
    return sums


def metrics_from_sums(sums: torch.Tensor, n_users: int, topk, metrics=("recall", "mrr", "ndcg", "hit", "precision"),
                      decimals: int | None = 4) -> OrderedDict:
    """The Evaluator.evaluate() dictionary: {'recall@10': ..} (base_metric.py:86-99)."""
    host = sums.cpu().numpy() / float(n_users)
    out = OrderedDict()
    for name in metrics:
        row = host[METRIC_ORDER.index(name.lower())]
        for k in topk:
            v = float(row[k - 1])
            out[f"{name.lower()}@{k}"] = round(v, decimals) if decimals is not None else v
    return out


class FusedCollector:
    """Collector twin fed by the fused top-k instead of dense scores.

    ``eval_batch_collect(model, user_ids, history_index, positive_u, positive_i)`` takes the same
    per-batch tuple the reference loader yields; ``get_data_struct()`` returns
    ``{"rec.topk": int32 [n_users, max(topk)+1], "rec.items": ids, "topk": topk}``.
    """

    def __init__(self, config):
        self.topk = list(config["topk"])
        self.kmax = max(self.topk)
        self._topk, self._items = [], []

    def eval_batch_collect(self, model, user_ids, history_index, positive_u, positive_i):
        device = next(model.parameters()).device
        users = torch.as_tensor(user_ids).to(device)
        n = users.numel()
        hist_off = hist_items = None
        if history_index is not None:
            hist_off, hist_items = csr_from_pairs(history_index[0], history_index[1], n, device)
        ids, _ = model.full_sort_topk(users, self.kmax, hist_off, hist_items, mask_pad=True, return_scores=False)
        pos_off, pos_items = csr_from_pairs(positive_u, positive_i, n, device)
        self._topk.append(topk_hits(ids, pos_off, pos_items))
        self._items.append(ids)
        return ids

    def get_data_struct(self):
        rec = torch.cat(self._topk) if self._topk else torch.zeros(0, self.kmax + 1, dtype=torch.int32)
        items = torch.cat(self._items) if self._items else torch.zeros(0, self.kmax, dtype=torch.int64)
        self._topk, self._items = [], []
        return {"rec.topk": rec, "rec.items": items, "topk": self.topk}


def evaluate_full_sort(model, user_ids, hist_off, hist_items, pos_off, pos_items, topk=(10,), decimals=4,
                       user_block: int | None = None, return_struct: bool = False):
    """Whole-split evaluation with device-resident CSR inputs; returns the metric dictionary."""
    device = next(model.parameters()).device
    users = torch.as_tensor(user_ids).to(device)
    n = users.numel()
    kmax = max(topk)
    block = user_block or n
    sums = torch.zeros(5, kmax, dtype=torch.float64, device=device)
    recs = []
    for s in range(0, n, block):
        e = min(n, s + block)
        ho = hi = None
        if hist_off is not None:
            base = hist_off[s]
            ho = (hist_off[s : e + 1] - base).contiguous()
            hi = hist_items[int(base.item()) : int(hist_off[e].item())] if block < n else hist_items
        ids, _ = model.full_sort_topk(users[s:e], kmax, ho, hi, mask_pad=True, return_scores=False)
        pbase = pos_off[s]
        po = (pos_off[s : e + 1] - pbase).contiguous()
        pi = pos_items[int(pbase.item()) : int(pos_off[e].item())] if block < n else pos_items
        rec = topk_hits(ids, po, pi)
        sums += topk_metric_sums(rec)
        if return_struct:
            recs.append(rec)
    result = metrics_from_sums(sums, n, topk, decimals=decimals)
    if return_struct:
        return result, torch.cat(recs)
    return result
