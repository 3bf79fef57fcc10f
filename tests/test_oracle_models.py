"""Pin the torch-CPU and float64 oracles to the reference's own outputs (tests/golden)."""

import numpy as np
import pytest
import torch

from oracle import kge_numpy as knp
from oracle.kge_torch import TABLE_NAMES, OracleKGE, Shapes, make_optimizer, train_step

from conftest import MODELS, load_golden


def _oracle_from_golden(name, g):
    U, I, E, R, d = (int(x) for x in g["shape"])
    torch.manual_seed(2024)
    m = OracleKGE(name, Shapes(U, I, E, R, d, margin=float(g["margin"])))
    return m, (U, I, E, R, d)


def _batch(g, i):
    keys = ("user_id", "item_id", "neg_item_id", "head_id", "relation_id", "tail_id", "neg_tail_id")
    return {k: torch.from_numpy(g[f"batch{i}/{k}"]) for k in keys}


@pytest.mark.parametrize("tag", ["d20", "d10"])
@pytest.mark.parametrize("name", MODELS)
def test_seeded_init_matches_reference(name, tag):
    g = load_golden(f"model_{name}_{tag}.npz")
    m, _ = _oracle_from_golden(name, g)
    sd = m.state_dict()
    assert sorted(sd) == sorted(k[5:] for k in g.files if k.startswith("init/"))
    for k, v in sd.items():
        np.testing.assert_array_equal(v.numpy(), g["init/" + k])


@pytest.mark.parametrize("tag", ["d20", "d10"])
@pytest.mark.parametrize("name", MODELS)
def test_training_trajectory_matches_reference(name, tag):
    g = load_golden(f"model_{name}_{tag}.npz")
    m, _ = _oracle_from_golden(name, g)
    opt = make_optimizer(m)
    for step, bi in enumerate(g["schedule"], start=1):
        loss = train_step(m, opt, _batch(g, int(bi)))
        np.testing.assert_allclose(loss, g["losses"][step - 1], rtol=1e-6)
        if f"step{step}/" + next(iter(m.state_dict())) in g.files:
            for k, v in m.state_dict().items():
                np.testing.assert_allclose(v.numpy(), g[f"step{step}/{k}"], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("tag", ["d20", "d10"])
@pytest.mark.parametrize("name", MODELS)
def test_scores_match_reference(name, tag):
    g = load_golden(f"model_{name}_{tag}.npz")
    m, _ = _oracle_from_golden(name, g)
    # the reference's scores were taken after its 12 training steps
    m.load_state_dict({k: torch.from_numpy(g["step12/" + k]) for k in m.state_dict()})
    b = _batch(g, 0)
    with torch.no_grad():
        np.testing.assert_allclose(m.predict(b).numpy(), g["predict"], rtol=1e-6, atol=1e-7)
        fs = m.full_sort_predict({"user_id": torch.from_numpy(g["fullsort_users"])})
        np.testing.assert_allclose(fs.numpy(), g["fullsort"], rtol=1e-6, atol=1e-7)
        if "predict_kg" in g.files:   # (transh.py scores users against items only)
            np.testing.assert_allclose(m.predict_kg(b).numpy(), g["predict_kg"], rtol=1e-6, atol=1e-7)
        if "fullsort_kg" in g.files:  # (transd.py:192-217 projects the head with <h, h>: not mirrored)
            kb = {"head_id": b["head_id"][:5], "relation_id": b["relation_id"][:5]}
            np.testing.assert_allclose(m.full_sort_predict_kg(kb).numpy(), g["fullsort_kg"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", MODELS)
def test_closed_form_gradients_match_reference(name):
    """The float64 loss/gradient closed forms (what the CUDA kernels implement) against the
    reference's autograd gradients at step 1."""
    g = load_golden(f"model_{name}_d20.npz")
    U, I, E, R, d = (int(x) for x in g["shape"])
    un, en, rn = TABLE_NAMES[name]
    tabs = {k[5:-7]: g[k].astype(np.float64) for k in g.files if k.startswith("init/")}
    b = {k: g[f"batch1/{k}"] for k in ("user_id", "item_id", "neg_item_id", "head_id", "relation_id", "tail_id", "neg_tail_id")}
    n_rec, n_kg = len(b["user_id"]), len(b["head_id"])
    w_rec, w_kg = knp.loss_weights(name, n_rec, n_kg)
    grads = {k: np.zeros_like(v) for k, v in tabs.items()}
    ui = np.full(n_rec, R - 1)
    total = 0.0
    segs = (
        (un, b["user_id"], ui, b["item_id"], b["neg_item_id"], w_rec),
        (en, b["head_id"], b["relation_id"], b["tail_id"], b["neg_tail_id"], w_kg),
    )
    for hnames, hid, rid, tpid, tnid, w in segs:
        h = [tabs[n][hid] for n in hnames]
        r = [tabs[n][rid] for n in rn]
        tp = [tabs[n][tpid] for n in en]
        tn = [tabs[n][tnid] for n in en]
        loss, gh, gr, gtp, gtn = knp.pair_loss_and_grads(name, h, r, tp, tn, w, margin=1.0)
        total += loss
        for n, gg in zip(hnames, gh):
            np.add.at(grads[n], hid, gg)
        for n, gg in zip(rn, gr):
            np.add.at(grads[n], rid, gg)
        for n, gg in zip(en, gtp):
            np.add.at(grads[n], tpid, gg)
        for n, gg in zip(en, gtn):
            np.add.at(grads[n], tnid, gg)
    # step 1 of the schedule uses batch 1
    assert int(g["schedule"][0]) == 1
    np.testing.assert_allclose(total, g["losses"][0], rtol=2e-6)
    for n, gg in grads.items():
        np.testing.assert_allclose(gg, g[f"grad1/{n}.weight"], rtol=2e-4, atol=2e-7)


def test_dense_adam_closed_form():
    rng = np.random.default_rng(0)
    p = rng.standard_normal((7, 5))
    tp = torch.nn.Parameter(torch.tensor(p, dtype=torch.float64))
    opt = torch.optim.Adam([tp], lr=1e-3)
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    for step in range(1, 30):
        g = rng.standard_normal(p.shape) * (rng.random((7, 1)) < 0.4)  # rows idle at random
        tp.grad = torch.tensor(g)
        opt.step()
        p, m, v = knp.adam_dense_step(p, m, v, g, step)
        np.testing.assert_allclose(p, tp.detach().numpy(), rtol=1e-12, atol=1e-14)
