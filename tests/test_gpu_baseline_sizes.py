"""Parity at the sizes BASELINE.json names (the small-shape tests pin the arithmetic; these pin it where the
benchmarks run).  The oracle (torch CPU, the reference's operators) is the checker; every case is sized so that
its CPU side finishes in seconds.

  cfg2  TransE d=100, E=30,001, the roofline batch 262,144 + 262,144: loss AND post-step weights of every table
  cfg3  RotatE d=256, 64 negatives in the reference's tiled layout (the positives repeated K times,
        abstract_dataloader.py:192-198), E reduced to 9,001: loss + weights after 3 steps
  cfg5  TransE d=128, E=1,000,001, 2,048 + 2,048: loss + the weights of every row the step touched, and no other
        row moved (dense Adam on the 1M-row tables runs on the host: ~2 s)
  cfg4  DistMult / ComplEx d=64, 200,001 items, top-20 through the TENSOR-CORE path: ids against the oracle's dense
        fp32 scores + the trainer's masking on 96 users (a rank flip is only accepted between items whose oracle
        scores differ by fp32 rounding of the two summation orders)
"""

import numpy as np
import pytest
import torch

from kge_helpers import (assert_weights_close, make_oracle_model, make_product_model, random_batch, tile_batch,
                         to_cpu_batch, to_device_batch)
from oracle import fullsort as ofs
from oracle.kge_torch import make_optimizer, train_step

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def test_cfg2_full_batch_loss_and_post_step_weights():
    U, I, E, R, d = 6041, 3001, 30001, 22, 100
    m = make_product_model("TransE", U, I, E, R, d)
    ora = make_oracle_model("TransE", U, I, E, R, d)
    opt = make_optimizer(ora)
    rng = np.random.default_rng(2024)
    for step in range(2):
        b = random_batch(rng, U, I, E, R, 262144, 262144)
        want = train_step(ora, opt, to_cpu_batch(b))
        loss = m.calculate_loss(to_device_batch(b))
        loss.backward()
        np.testing.assert_allclose(loss.item(), want, rtol=RTOL, err_msg=f"loss step {step + 1}")
    for (k, vo), (_, vp) in zip(ora.state_dict().items(), m.state_dict().items()):
        # every row collects ~40 gradient rows per step here: the fp32 sum order of the atomics differs from
        # autograd's index_add, which Adam's 1/sqrt(v) amplifies on near-cancelling elements (kge_helpers)
        assert_weights_close(vp.cpu().numpy(), vo.numpy(), rtol=RTOL, atol=1e-6, outlier_frac=2e-3, err_msg=k)


def test_cfg3_rotate_64_negatives_tiled_layout():
    U, I, E, R, d, K = 4592, 4554, 9001, 44, 256, 64
    m = make_product_model("RotatE", U, I, E, R, d)
    ora = make_oracle_model("RotatE", U, I, E, R, d)
    opt = make_optimizer(ora)
    rng = np.random.default_rng(3)
    for step in range(3):
        b = tile_batch(random_batch(rng, U, I, E, R, 2048, 2048, K, K), K, K)   # what the reference loader yields
        assert b["user_id"].shape[0] == 2048 * K and b["neg_tail_id"].shape[0] == 2048 * K
        want = train_step(ora, opt, to_cpu_batch(b))
        loss = m.calculate_loss(to_device_batch(b))
        loss.backward()
        np.testing.assert_allclose(loss.item(), want, rtol=RTOL, err_msg=f"loss step {step + 1}")
    for (k, vo), (_, vp) in zip(ora.state_dict().items(), m.state_dict().items()):
        assert_weights_close(vp.cpu().numpy(), vo.numpy(), rtol=RTOL, atol=1e-6, outlier_frac=2e-3, err_msg=k)


def test_cfg5_million_entity_step_touched_rows():
    U, I, E, R, d = 115001, 30001, 1000001, 54, 128
    m = make_product_model("TransE", U, I, E, R, d)
    ora = make_oracle_model("TransE", U, I, E, R, d)
    opt = make_optimizer(ora)
    before = m.entity_embedding.weight.detach().clone()
    rng = np.random.default_rng(5)
    b = random_batch(rng, U, I, E, R, 2048, 2048)
    want = train_step(ora, opt, to_cpu_batch(b))
    loss = m.calculate_loss(to_device_batch(b))
    loss.backward()
    np.testing.assert_allclose(loss.item(), want, rtol=RTOL)
    sd_o, sd_p = ora.state_dict(), m.state_dict()
    ent = np.unique(np.concatenate([b["item_id"], b["neg_item_id"], b["head_id"], b["tail_id"], b["neg_tail_id"]]))
    usr = np.unique(b["user_id"])
    rel = np.unique(np.concatenate([b["relation_id"], [R - 1]]))
    for key, rows in (("entity_embedding.weight", ent), ("user_embedding.weight", usr),
                      ("relation_embedding.weight", rel)):
        got, ref = sd_p[key].cpu().numpy(), sd_o[key].numpy()
        assert_weights_close(got[rows], ref[rows], rtol=RTOL, atol=5e-7, err_msg=key)
        # rows the step did not touch: first step, so dense Adam leaves them exactly where they were
        mask = np.ones(got.shape[0], dtype=bool)
        mask[rows] = False
        assert np.array_equal(got[mask], ref[mask]), key
    assert torch.equal(m.entity_embedding.weight.detach()[~torch.from_numpy(np.isin(np.arange(E), ent)).cuda()],
                       before[~torch.from_numpy(np.isin(np.arange(E), ent)).cuda()])


@pytest.mark.parametrize("name", ["DistMult", "ComplEx"])
def test_cfg4_tensor_core_topk_against_oracle_scores(name):
    U, I, d, k, n = 5001, 200001, 64, 20, 96
    m = make_product_model(name, U, I, I, 3, d)
    ora = make_oracle_model(name, U, I, I, 3, d)
    rng = np.random.default_rng(17)
    users = rng.integers(1, U, n)
    hist = np.sort(rng.integers(1, I, (n, 50)), axis=1)
    off = torch.arange(0, 50 * n + 1, 50, dtype=torch.long).cuda()
    ids, sc = m.full_sort_topk(torch.from_numpy(users).cuda(), k, off, torch.from_numpy(hist.reshape(-1)).cuda(),
                               path="mma")
    assert m._mma_last_fallback_rows == 0
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    with torch.no_grad():   # (8 users at a time: the reference forms [users, items, d] temporaries)
        dense = np.concatenate([ora.full_sort_predict({"user_id": torch.from_numpy(users[s:s + 8])}).view(-1, I).numpy()
                                for s in range(0, n, 8)]).copy()
    dense[:, 0] = -np.inf                                   # trainer.py:731-734
    for r in range(n):
        dense[r, hist[r]] = -np.inf
    flips = 0
    for r in range(n):
        want_ids, want_sc = ofs.topk_canonical(dense[r : r + 1], k)
        want_ids, want_sc = want_ids[0], want_sc[0]
        np.testing.assert_allclose(sc[r], dense[r, ids[r]], rtol=RTOL, atol=1e-7)   # reported scores are the fp32 scores
        if not np.array_equal(ids[r], want_ids):
            # same set up to near-ties: every differing item's oracle score is within rounding of the k-th score
            tau = want_sc[-1]
            scale = np.abs(want_sc).max()
            for j in set(ids[r]) ^ set(want_ids):
                assert abs(dense[r, j] - tau) <= 4e-6 * scale, (r, j, dense[r, j], tau)
            flips += 1
    assert flips <= 2, f"{flips} of {n} rows differ from the oracle's order beyond near-ties"
