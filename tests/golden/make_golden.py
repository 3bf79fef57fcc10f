"""Generate the golden fixtures from the UNMODIFIED reference (hopwise @ /root/reference).

Run in the build container only:   python tests/golden/make_golden.py
Writes tests/golden/*.npz.  The fixtures are what pins the oracle (oracle/) and, through
the `-m gpu` tests, the CUDA path to the reference's own outputs.

Nothing here is copied from the reference; it only calls its public classes:
  hopwise.model.knowledge_graph_embedding_recommender.{transe,rotate,distmult,complex}
  hopwise.sampler.{KGSampler,Sampler}
  hopwise.evaluator.{Collector,Evaluator} and metrics
"""

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_harness import REF_CONFIG, FakeDataset, import_reference  # noqa: E402

import_reference()
from hopwise.data.interaction import Interaction  # noqa: E402
from hopwise.evaluator import Collector, Evaluator  # noqa: E402
from hopwise.model.knowledge_graph_embedding_recommender.complex import ComplEx  # noqa: E402
from hopwise.model.knowledge_graph_embedding_recommender.distmult import DistMult  # noqa: E402
from hopwise.model.knowledge_graph_embedding_recommender.rotate import RotatE  # noqa: E402
from hopwise.model.knowledge_graph_embedding_recommender.toruse import TorusE  # noqa: E402
from hopwise.model.knowledge_graph_embedding_recommender.transd import TransD  # noqa: E402
from hopwise.model.knowledge_graph_embedding_recommender.transe import TransE  # noqa: E402
from hopwise.model.knowledge_graph_embedding_recommender.transh import TransH  # noqa: E402
from hopwise.sampler import KGSampler, Sampler  # noqa: E402

MODELS = {"TransE": TransE, "RotatE": RotatE, "DistMult": DistMult, "ComplEx": ComplEx, "TorusE": TorusE,
          "TransH": TransH, "TransD": TransD}
SEED = 2024
# (n_users, n_items, n_entities, n_relations, d)
SHAPES = {"d20": (37, 23, 61, 7, 20), "d10": (19, 11, 29, 5, 10)}
N_STEPS = 12
SAVE_AT = (1, 4, 12)


def make_batches(rng, U, I, E, R, n_batches, n_rec, n_kg):
    out = []
    for _ in range(n_batches):
        out.append(
            {
                "user_id": rng.integers(1, U, n_rec),
                "item_id": rng.integers(1, I, n_rec),
                "neg_item_id": rng.integers(1, I, n_rec),
                "head_id": rng.integers(1, E, n_kg),
                "relation_id": rng.integers(1, R - 1, n_kg),
                "tail_id": rng.integers(1, E, n_kg),
                "neg_tail_id": rng.integers(1, E, n_kg),
            }
        )
    return out


def to_inter(b):
    return Interaction({k: torch.as_tensor(v, dtype=torch.long) for k, v in b.items()})


def golden_models(only=None):
    for tag, (U, I, E, R, d) in SHAPES.items():
        rng = np.random.default_rng(SEED)
        # ragged rec/KG halves, duplicates guaranteed by the small id ranges; only a few
        # distinct batches so that some rows sit untouched for several Adam steps
        batches = make_batches(rng, U, I, E, R, 3, n_rec=13, n_kg=17)
        for name, cls in MODELS.items():
            if only is not None and name not in only:
                continue
            cfg = dict(REF_CONFIG, embedding_size=d, margin=1.0)
            ds = FakeDataset(U, I, E, R)
            torch.manual_seed(SEED)
            model = cls(cfg, ds)
            out = {"shape": np.array([U, I, E, R, d]), "margin": np.float32(1.0)}
            for k, v in model.state_dict().items():
                out["init/" + k] = v.numpy().copy()
            for bi, b in enumerate(batches):
                for k, v in b.items():
                    out[f"batch{bi}/{k}"] = v.astype(np.int64)
            opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0.0)
            losses = []
            for step in range(1, N_STEPS + 1):
                # batch schedule leaves batch 2's rows idle between steps 3 and 12
                b = batches[{3: 2, 12: 2}.get(step, step % 2)]
                opt.zero_grad()
                loss = model.calculate_loss(to_inter(b))
                losses.append(loss.item())
                loss.backward()
                if step == 1:
                    for k, p in model.named_parameters():
                        out["grad1/" + k] = p.grad.numpy().copy()
                opt.step()
                if step in SAVE_AT:
                    for k, v in model.state_dict().items():
                        out[f"step{step}/" + k] = v.numpy().copy()
            out["losses"] = np.array(losses, dtype=np.float64)
            out["schedule"] = np.array([{3: 2, 12: 2}.get(s, s % 2) for s in range(1, N_STEPS + 1)])
            with torch.no_grad():
                users = torch.arange(1, min(U, 9))
                pb = to_inter(batches[0])
                out["predict"] = model.predict(pb).numpy()
                has_kg = name != "TransH"   # transh.py scores users against items only
                if has_kg:
                    out["predict_kg"] = model.predict_kg(pb).numpy()
                out["fullsort_users"] = users.numpy()
                out["fullsort"] = model.full_sort_predict(Interaction({"user_id": users})).view(-1, I).numpy()
                if has_kg and name != "TransD":   # (transd.py:192-217 projects the head with <h, h>: not mirrored)
                    kb = Interaction({"head_id": pb["head_id"][:5], "relation_id": pb["relation_id"][:5]})
                    out["fullsort_kg"] = model.full_sort_predict_kg(kb).view(-1, E).numpy()
            np.savez_compressed(os.path.join(HERE, f"model_{name}_{tag}.npz"), **out)
            print("wrote", name, tag, "loss[0]", losses[0], "loss[-1]", losses[-1])


class _Feat(dict):
    pass


class _RecDataset:
    def __init__(self, users, items, n_users, n_items):
        self.uid_field, self.iid_field = "user_id", "item_id"
        self.user_num, self.item_num = n_users, n_items
        self.inter_feat = {"user_id": torch.as_tensor(users), "item_id": torch.as_tensor(items)}


def golden_sampler():
    rng = np.random.default_rng(SEED)
    E, U, I = 61, 37, 23
    n_tri = 400
    heads = rng.integers(1, E, n_tri)
    tails = rng.integers(1, E, n_tri)
    # one hub head linked to most entities: stresses multi-round rejection
    hub_tails = rng.permutation(np.arange(1, E))[: E - 6]
    heads = np.concatenate([heads, np.full(len(hub_tails), 7)])
    tails = np.concatenate([tails, hub_tails])
    ds = FakeDataset(U, I, E, 7, heads=heads, tails=tails)
    kg = KGSampler(ds)
    ru = rng.integers(1, U, 300)
    ri = rng.integers(1, I, 300)
    rec = Sampler("train", _RecDataset(ru, ri, U, I)).set_phase("train")

    out = {"E": E, "U": U, "I": I, "heads": heads, "tails": tails, "rec_users": ru, "rec_items": ri}
    np.random.seed(SEED)
    out["state0_key"], out["state0_pos"] = np.random.get_state()[1].copy(), np.random.get_state()[2]
    calls = []
    q_heads = [heads[rng.integers(0, len(heads), 40)], np.full(9, 7), heads[rng.integers(0, len(heads), 25)]]
    q_users = [ru[rng.integers(0, len(ru), 31)], ru[rng.integers(0, len(ru), 8)], np.full(5, ru[0])]
    nums = [1, 3, 2]
    # the trainer's order per step: KG negatives first, then rec negatives, one shared stream
    for step, (qh, qu, num) in enumerate(zip(q_heads, q_users, nums)):
        neg_t = kg.sample_by_entity_ids(qh, num).numpy()
        st = np.random.get_state()
        out[f"call{step}/heads"], out[f"call{step}/num"] = qh, num
        out[f"call{step}/neg_tails"] = neg_t
        out[f"call{step}/kg_key"], out[f"call{step}/kg_pos"] = st[1].copy(), st[2]
        neg_i = rec.sample_by_user_ids(qu, None, num).numpy()
        st = np.random.get_state()
        out[f"call{step}/users"] = qu
        out[f"call{step}/neg_items"] = neg_i
        out[f"call{step}/rec_key"], out[f"call{step}/rec_pos"] = st[1].copy(), st[2]
        calls.append(step)
    out["n_calls"] = len(calls)
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **out)
    print("wrote sampler", [len(out[f'call{c}/neg_tails']) for c in calls])


def golden_sampler_pop():
    """Popularity-biased negatives (sampler.py:68-116) for the KG and the rec sampler, alpha 1.0 and 0.5, sharing one
    stream; includes the alias table itself so the table build is pinned independently of the draws."""
    rng = np.random.default_rng(SEED + 1)
    E, U, I = 53, 29, 31
    n_tri = 500
    # skewed popularity: squares of uniforms concentrate on small ids
    heads = 1 + (rng.random(n_tri) ** 2 * (E - 1)).astype(np.int64)
    tails = 1 + (rng.random(n_tri) ** 3 * (E - 1)).astype(np.int64)
    ru = rng.integers(1, U, 400)
    ri = 1 + (rng.random(400) ** 2 * (I - 1)).astype(np.int64)
    out = {"E": E, "U": U, "I": I, "heads": heads, "tails": tails, "rec_users": ru, "rec_items": ri}
    for tag, alpha in (("a1", 1.0), ("a05", 0.5)):
        ds = FakeDataset(U, I, E, 7, heads=heads, tails=tails)
        kg = KGSampler(ds, distribution="popularity", alpha=alpha)
        rec = Sampler("train", _RecDataset(ru, ri, U, I), distribution="popularity", alpha=alpha).set_phase("train")
        for name, smp in (("kg", kg), ("rec", rec)):
            keys = list(smp.prob.keys())
            out[f"{tag}/{name}_keys"] = np.array(keys, dtype=np.int64)
            out[f"{tag}/{name}_prob"] = np.array([smp.prob[k] for k in keys], dtype=np.float64)
            out[f"{tag}/{name}_alias"] = np.array([smp.alias[k] for k in keys], dtype=np.int64)
        np.random.seed(SEED + 7)
        # an odd number of words consumed first: the doubles' word pairs then straddle the 624-word blocks
        np.random.randint(0, 1 << 30, 3)
        st = np.random.get_state()
        out[f"{tag}/state0_key"], out[f"{tag}/state0_pos"] = st[1].copy(), st[2]
        q_heads = [heads[rng.integers(0, len(heads), 350)], np.full(11, heads[0]), heads[rng.integers(0, len(heads), 27)]]
        q_users = [ru[rng.integers(0, len(ru), 330)], ru[rng.integers(0, len(ru), 9)], np.full(6, ru[0])]
        nums = [2, 3, 1]
        for step, (qh, qu, num) in enumerate(zip(q_heads, q_users, nums)):
            out[f"{tag}/call{step}/heads"], out[f"{tag}/call{step}/users"], out[f"{tag}/call{step}/num"] = qh, qu, num
            out[f"{tag}/call{step}/neg_tails"] = kg.sample_by_entity_ids(qh, num).numpy()
            st = np.random.get_state()
            out[f"{tag}/call{step}/kg_key"], out[f"{tag}/call{step}/kg_pos"] = st[1].copy(), st[2]
            out[f"{tag}/call{step}/neg_items"] = rec.sample_by_user_ids(qu, None, num).numpy()
            st = np.random.get_state()
            out[f"{tag}/call{step}/rec_key"], out[f"{tag}/call{step}/rec_pos"] = st[1].copy(), st[2]
        out[f"{tag}/n_calls"] = len(nums)
    np.savez_compressed(os.path.join(HERE, "sampler_pop.npz"), **out)
    print("wrote sampler_pop", [len(out[f"a1/call{c}/neg_tails"]) for c in range(3)])


def golden_eval():
    """Collector + Evaluator on masked random scores (no ties => torch.topk is canonical)."""
    rng = np.random.default_rng(SEED)
    n, I, k = 40, 57, 10
    scores = rng.standard_normal((n, I)).astype(np.float32)
    hist_u, hist_i, pos_u, pos_i = [], [], [], []
    for u in range(n):
        perm = rng.permutation(np.arange(1, I))
        nh, npos = rng.integers(0, 30), rng.integers(1, 6)
        if u == 3:
            nh = I - 1 - 4  # fewer than k unmasked items: -inf enters the top-k
            npos = 2
        hist_u += [u] * nh
        hist_i += list(perm[:nh])
        pos_u += [u] * npos
        pos_i += list(perm[nh : nh + npos])
    cfg = {
        "eval_args": {"mode": {"valid": "full", "test": "full"}},
        "topk": [5, k],
        "device": torch.device("cpu"),
        "metrics": ["Recall", "MRR", "NDCG", "Hit", "Precision"],
        "metric_decimal_place": 4,
        "tsne": None,
    }
    cfg["eval_args"]["mode"] = "full"
    coll = Collector(cfg)
    s = torch.from_numpy(scores.copy())
    s[:, 0] = -np.inf
    s[(torch.as_tensor(hist_u), torch.as_tensor(hist_i))] = -np.inf
    coll.eval_batch_collect(s, None, torch.as_tensor(pos_u), torch.as_tensor(pos_i))
    struct = coll.get_data_struct()
    rec_topk = struct.get("rec.topk").numpy()
    result = Evaluator(cfg).evaluate(struct)
    _, ids = torch.topk(s, k, dim=-1)
    out = {
        "scores": scores,
        "hist_u": np.array(hist_u),
        "hist_i": np.array(hist_i),
        "pos_u": np.array(pos_u),
        "pos_i": np.array(pos_i),
        "k": k,
        "topk_ids": ids.numpy(),
        "rec_topk": rec_topk,
        "metric_names": np.array(list(result.keys())),
        "metric_values": np.array(list(result.values()), dtype=np.float64),
    }
    # unrounded means straight from the metric classes
    from hopwise.evaluator.register import metrics_dict

    pos_index = rec_topk[:, :k].astype(bool)
    pos_len = rec_topk[:, k]
    for name in ("recall", "mrr", "ndcg", "hit", "precision"):
        m = metrics_dict[name](cfg)
        mat = m.metric_info(pos_index, pos_len) if name in ("recall", "ndcg") else m.metric_info(pos_index)
        out["matrix/" + name] = np.asarray(mat, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "eval.npz"), **out)
    print("wrote eval", dict(result))


if __name__ == "__main__":
    torch.set_num_threads(1)
    parts = {"models": golden_models, "sampler": golden_sampler, "sampler_pop": golden_sampler_pop, "eval": golden_eval,
             # (models added later: the others stay untouched)
             "toruse": lambda: golden_models(only=("TorusE",)), "transh": lambda: golden_models(only=("TransH",)),
             "transd": lambda: golden_models(only=("TransD",))}
    for name in sys.argv[1:] or [p for p in parts if p not in ("toruse", "transh", "transd")]:   # e.g. `make_golden.py sampler_pop`
        parts[name]()
