"""Torch-CPU restatement of hopwise's four KGE recommenders (oracle; test infrastructure only).

One spec-driven class replaces the reference's four files.  It uses the same torch
operators the reference calls, so on the CPU it reproduces the reference's fp32
results (the golden fixtures under tests/golden pin that).

Reference call sites followed (relative to /root/reference/hopwise/model/
knowledge_graph_embedding_recommender/):
  transe.py:36-53 (tables, TripletMarginLoss p=2), :55-57 (score), :59-98 (loss),
      :100-126 (predict / full_sort_predict), :128-154 (kg twins)
  distmult.py:32-48, :50-51, :53-95, :97-133, :110-146
  rotate.py:36-59 (phase table re-initialised U(0, 2pi)), :61-69 (single L2 norm over
      the stacked (re, im) residual), :98-131 (two BCE means), :133-190
  complex.py:31-51, :53-62 (4th term uses tail_im, as written), :95-128, :130-190
  ../init.py:13-29 (xavier_normal_ on every nn.Embedding)
  ../abstract_recommender.py:204-223 (field names, n_users/n_items/n_entities/n_relations)
Optimiser: torch.optim.Adam(lr, weight_decay) as built in trainer/trainer.py:189-190;
step order as in trainer/trainer.py:243-265.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import torch
from torch import nn

# table-name layout per model: (user parts, entity parts, relation parts).  Creation
# order (users, entities, relations; re before im) is the reference's constructor
# order, which fixes the RNG stream of the initialisation.
TABLE_NAMES = {
    "TransE": (["user_embedding"], ["entity_embedding"], ["relation_embedding"]),
    "DistMult": (["user_embedding"], ["entity_embedding"], ["relation_embedding"]),
    # toruse.py:31-48: TransE's tables
    "TorusE": (["user_embedding"], ["entity_embedding"], ["relation_embedding"]),
    # transh.py:31-48: the second relation table is the hyperplane vector
    "TransH": (["user_embedding"], ["entity_embedding"], ["relation_embedding", "norm_vec"]),
    # transd.py:41-47: an embedding and a transfer vector per table (created embeddings first, then the vectors)
    "TransD": (["user_embedding", "user_vec_embedding"], ["entity_embedding", "entity_vec_embedding"],
               ["relation_embedding", "relation_vec_embedding"]),
    "RotatE": (
        ["user_embedding", "user_embedding_im"],
        ["entity_embedding", "entity_embedding_im"],
        ["relation_embedding"],
    ),
    "ComplEx": (
        ["user_re_embedding", "user_im_embedding"],
        ["entity_re_embedding", "entity_im_embedding"],
        ["relation_re_embedding", "relation_im_embedding"],
    ),
}


@dataclass
class Shapes:
    n_users: int
    n_items: int
    n_entities: int
    n_relations: int
    embedding_size: int
    margin: float = 1.0
    ui_relation: int | None = None  # token id of [UI-Relation]; default = last row

    def __post_init__(self):
        if self.ui_relation is None:
            self.ui_relation = self.n_relations - 1


class OracleKGE(nn.Module):
    """CPU oracle with the KnowledgeRecommender call surface (tensor-dict batches)."""

    def __init__(self, model: str, shapes: Shapes):
        super().__init__()
        if model not in TABLE_NAMES:
            raise ValueError(model)
        self.model = model
        self.shapes = shapes
        d = shapes.embedding_size
        unames, enames, rnames = TABLE_NAMES[model]
        rows = {**{n: shapes.n_users for n in unames}, **{n: shapes.n_entities for n in enames},
                **{n: shapes.n_relations for n in rnames}}
        order = unames + enames + rnames
        if model == "TransD":   # the reference constructor's order (it fixes the RNG stream of the initialisation)
            order = [unames[0], enames[0], rnames[0], unames[1], enames[1], rnames[1]]
        for n in order:
            setattr(self, n, nn.Embedding(rows[n], d))
        for mod in self.modules():  # same traversal order as nn.Module.apply on children
            if isinstance(mod, nn.Embedding):
                nn.init.xavier_normal_(mod.weight.data)
        if model == "RotatE":
            nn.init.uniform_(self.relation_embedding.weight, 0, 2 * math.pi)
        self._u = [getattr(self, n) for n in unames]
        self._e = [getattr(self, n) for n in enames]
        self._r = [getattr(self, n) for n in rnames]

    # ---- relation row used for user->item triples -------------------------------------
    def _ui_row(self, full_sort: bool) -> int:
        # TransE / DistMult always take weight[-1]; RotatE always the token id; ComplEx the
        # token id for loss/predict and weight[-1] for full_sort_predict.
        if self.model in ("TransE", "DistMult", "TorusE", "TransD"):   # toruse.py:53, 121, 131; transd.py:60, 69
            return self.shapes.n_relations - 1
        if self.model == "TransH":
            # transh.py:63, 90: relation_embedding.weight[-1] but norm_vec(ui_relation) -- one row only when the token
            # is the last relation id, which is the case the product supports
            assert self.shapes.ui_relation == self.shapes.n_relations - 1
            return self.shapes.n_relations - 1
        if self.model == "ComplEx" and full_sort:
            return self.shapes.n_relations - 1
        return self.shapes.ui_relation

    # ---- scorers (broadcast over leading dims; reduce the last) -----------------------
    def score(self, h, r, t):
        m = self.model
        if m == "TransE":
            return -torch.norm(h[0] + r[0] - t[0], p=2, dim=-1)
        if m == "TransD":   # transd.py:86-91, 135-176: proj(e) = r_p * <e, e_p> + e on head and tail, then TransE's norm
            proj = lambda e: r[1] * (e[0] * e[1]).sum(dim=-1, keepdim=True) + e[0]   # noqa: E731
            return -torch.norm(proj(h) + r[0] - proj(t), p=2, dim=-1)
        if m == "TransH":   # transh.py:53-58, 73-74: project(e) = e - (e * w.sum()) * w
            w = r[1]
            sw = w.sum(dim=-1, keepdim=True)
            proj = lambda e: e - (e * sw) * w   # noqa: E731
            return -torch.norm(proj(h[0]) + r[0] - proj(t[0]), p=2, dim=-1)
        if m == "TorusE":   # toruse.py:66-76 (frac_ on detached copies: x - trunc(x), sign kept)
            x = (torch.frac(h[0]) + torch.frac(r[0])) - torch.frac(t[0])
            return -(4 * torch.min(x ** 2, 1 - x ** 2).sum(dim=-1))
        if m == "DistMult":
            return (h[0] * r[0] * t[0]).sum(dim=-1)
        if m == "RotatE":
            c, s = torch.cos(r[0]), torch.sin(r[0])
            re = (c * h[0] - s * h[1]) - t[0]
            im = (c * h[1] + s * h[0]) - t[1]
            both = torch.stack([re, im], dim=-1)
            return self.shapes.margin - torch.linalg.vector_norm(both, dim=(-2, -1))
        # ComplEx, with the reference's fourth term on tail_im
        hr, hi = h
        rr, ri = r
        tr, ti = t
        return (
            (hr * rr * tr).sum(-1) + (hi * rr * ti).sum(-1) + (hr * ri * ti).sum(-1) - (hi * ri * ti).sum(-1)
        )

    def _rows(self, tabs, idx):
        return [tab(idx) for tab in tabs]

    # ---- training loss ----------------------------------------------------------------
    def calculate_loss(self, batch: dict) -> torch.Tensor:
        user, item, neg_item = batch["user_id"], batch["item_id"], batch["neg_item_id"]
        head, rel = batch["head_id"], batch["relation_id"]
        tail, neg_tail = batch["tail_id"], batch["neg_tail_id"]
        ui = torch.full_like(user, self._ui_row(full_sort=False))

        u = self._rows(self._u, user)
        ur = self._rows(self._r, ui)
        ip, ineg = self._rows(self._e, item), self._rows(self._e, neg_item)
        h = self._rows(self._e, head)
        kr = self._rows(self._r, rel)
        tp, tn = self._rows(self._e, tail), self._rows(self._e, neg_tail)

        if self.model in ("TransE", "TorusE"):   # toruse.py:81-102 is transe.py:75-98: the torus only scores
            anchor = torch.cat([u[0] + ur[0], h[0] + kr[0]])
            pos, neg = torch.cat([ip[0], tp[0]]), torch.cat([ineg[0], tn[0]])
            return nn.functional.triplet_margin_loss(anchor, pos, neg, margin=self.shapes.margin, p=2)
        if self.model == "TransD":   # transd.py:93-133
            def projd(e, rr):
                return rr[1] * (e[0] * e[1]).sum(dim=1, keepdim=True) + e[0]

            anchor = torch.cat([projd(u, ur) + ur[0], projd(h, kr) + kr[0]])
            pos = torch.cat([projd(ip, ur), projd(tp, kr)])
            neg = torch.cat([projd(ineg, ur), projd(tn, kr)])
            return nn.functional.triplet_margin_loss(anchor, pos, neg, margin=self.shapes.margin, p=2)
        if self.model == "TransH":   # transh.py:76-107: project every row with its triple's relation, then TransE's loss
            def proj(e, rr):
                return e - (e * rr[1].sum(dim=1, keepdim=True)) * rr[1]

            anchor = torch.cat([proj(u[0], ur) + ur[0], proj(h[0], kr) + kr[0]])
            pos = torch.cat([proj(ip[0], ur), proj(tp[0], kr)])
            neg = torch.cat([proj(ineg[0], ur), proj(tn[0], kr)])
            return nn.functional.triplet_margin_loss(anchor, pos, neg, margin=self.shapes.margin, p=2)
        if self.model == "DistMult":
            cat = lambda a, b: [torch.cat([x, y]) for x, y in zip(a, b)]  # noqa: E731
            hh, rr = cat(u, h), cat(ur, kr)
            s_pos, s_neg = self.score(hh, rr, cat(ip, tp)), self.score(hh, rr, cat(ineg, tn))
            return nn.functional.margin_ranking_loss(
                s_pos, s_neg, torch.ones_like(s_pos), margin=self.shapes.margin
            )
        # RotatE / ComplEx: one BCE-with-logits mean per task, summed
        total = 0
        for hh, rr, p, n in ((u, ur, ip, ineg), (h, kr, tp, tn)):
            sp, sn = self.score(hh, rr, p), self.score(hh, rr, n)
            logits = torch.cat([sp, sn])
            labels = torch.cat([torch.ones_like(sp), torch.zeros_like(sn)])
            total = total + nn.functional.binary_cross_entropy_with_logits(logits, labels)
        return total

    # ---- inference ----------------------------------------------------------------------
    def predict(self, batch: dict) -> torch.Tensor:
        user, item = batch["user_id"], batch["item_id"]
        ui = torch.full_like(user, self._ui_row(full_sort=False))
        return self.score(self._rows(self._u, user), self._rows(self._r, ui), self._rows(self._e, item))

    def predict_kg(self, batch: dict) -> torch.Tensor:
        return self.score(
            self._rows(self._e, batch["head_id"]),
            self._rows(self._r, batch["relation_id"]),
            self._rows(self._e, batch["tail_id"]),
        )

    def full_sort_predict(self, batch: dict) -> torch.Tensor:
        user = batch["user_id"]
        ui = torch.full_like(user, self._ui_row(full_sort=True))
        items = torch.arange(self.shapes.n_items)
        h = [x.unsqueeze(1) for x in self._rows(self._u, user)]
        r = [x.unsqueeze(1) for x in self._rows(self._r, ui)]
        t = [x.unsqueeze(0) for x in self._rows(self._e, items)]
        return self.score(h, r, t)

    def full_sort_predict_kg(self, batch: dict) -> torch.Tensor:
        ents = torch.arange(self.shapes.n_entities)
        h = [x.unsqueeze(1) for x in self._rows(self._e, batch["head_id"])]
        r = [x.unsqueeze(1) for x in self._rows(self._r, batch["relation_id"])]
        t = [x.unsqueeze(0) for x in self._rows(self._e, ents)]
        return self.score(h, r, t)


def make_optimizer(model: nn.Module, lr: float = 1e-3, weight_decay: float = 0.0, learner: str = "adam"):
    """The optimiser the reference trainer builds for `learner` (trainer/trainer.py:189-205): torch's defaults apart
    from lr and weight_decay."""
    cls = {"adam": torch.optim.Adam, "adamw": torch.optim.AdamW, "sgd": torch.optim.SGD, "adagrad": torch.optim.Adagrad,
           "rmsprop": torch.optim.RMSprop}[learner.lower()]
    return cls(model.parameters(), lr=lr, weight_decay=weight_decay)


def train_step(model: OracleKGE, opt, batch: dict) -> float:
    """One optimisation step in the trainer's order (trainer.py:243-265)."""
    opt.zero_grad()
    loss = model.calculate_loss(batch)
    value = loss.item()
    loss.backward()
    opt.step()
    return value
