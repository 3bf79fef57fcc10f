"""Two-GPU NCCL test of the row-sparse data-parallel train step (skipped with fewer than 2 GPUs):
replicas stay bit-identical, and equal the single-GPU step on the concatenated batch (DDP averages
the gradients of per-rank mean losses = the global-batch mean for equal shards, trainer.py:82-112)."""

import os
import socket

import numpy as np
import pytest
import torch

from kge_helpers import assert_weights_close, make_product_model, random_batch, to_device_batch

pytestmark = pytest.mark.gpu

# small tables: every table takes the dense all-reduce route; large ones: user / entity lists are exchanged
SHAPES = {"dense": dict(U=300, I=200, E=700, R=9, d=64), "sparse": dict(U=3000, I=2500, E=9000, R=9, d=64)}
N_REC, N_KG, STEPS = 192, 160, 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batches(name, SHAPE):
    rng = np.random.default_rng(5)
    k = 3 if name in ("RotatE", "ComplEx") else 1
    return [random_batch(rng, SHAPE["U"], SHAPE["I"], SHAPE["E"], SHAPE["R"], 2 * N_REC, 2 * N_KG, k, k) for _ in range(STEPS)], k


def _shard(b, rank, k):
    """Rank's half of a global batch; negatives are j-major [k, n] (sampler.py:146-153)."""
    out = {}
    for key, v in b.items():
        if key.startswith("neg_"):
            n = v.shape[0] // k
            half = n // 2
            out[key] = v.reshape(k, n)[:, rank * half:(rank + 1) * half].reshape(-1)
        else:
            half = v.shape[0] // 2
            out[key] = v[rank * half:(rank + 1) * half]
    return out


def _worker(rank, world, port, name, route, out_dir, multimem=False, owner=False):
    import torch.distributed as dist

    from hopwise_b200.distributed import broadcast_weights, enable_row_sparse_data_parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dev = f"cuda:{rank}"
        SHAPE = SHAPES[route]
        m = make_product_model(name, device=dev, seed=2024 + rank, **SHAPE)   # different init: broadcast must fix it
        broadcast_weights(m)
        ex = enable_row_sparse_data_parallel(m, multimem=multimem, owner_adam=owner)
        batches, k = _batches(name, SHAPE)
        losses = []
        for b in batches:
            loss = m.calculate_loss(to_device_batch(_shard(b, rank, k), dev))
            loss.backward()
            losses.append(float(loss.item()))
        assert ex.bytes_per_step > 0
        if multimem:   # the in-switch reduction really ran (one launch of this library's kernel per step)
            assert ex.symm is not None and ex.kernels_per_step >= 1
        if owner:      # the parameters are views of the symmetric weight buffer; no row-lazy state is kept
            assert ex.owner_adam and m._state["w_flat"].data_ptr() == next(m.parameters()).data_ptr()
            assert m.kge_optimizer_state is None
        assert ex.dense == ([True, True, True] if route == "dense" or owner else [False, False, True])
        sd = {key: v.cpu().numpy() for key, v in m.state_dict().items()}
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), losses=np.array(losses), **sd)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,route,multimem", [("TransE", "dense", False), ("ComplEx", "dense", False),
                                                 ("TransE", "sparse", False), ("RotatE", "sparse", False),
                                                 ("TransE", "dense", True), ("ComplEx", "dense", True),
                                                 ("TransE", "dense", "owner"), ("RotatE", "dense", "owner"),
                                                 ("ComplEx", "sparse", "owner")])
def test_two_rank_row_sparse_step_matches_single_gpu(name, route, multimem, tmp_path):
    """multimem = the dense route reduced in the NVSwitch by kge_multimem_all_reduce_f32 instead of NCCL;
    "owner" = the whole optimiser step owner-sharded over the switch (kge_owner_adam_step; dense Adam on every table,
    whatever share of the rows the batch touches)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    SHAPE = SHAPES[route]
    owner = multimem == "owner"
    mp.spawn(_worker, args=(2, _free_port(), name, route, str(tmp_path), bool(multimem), owner), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    keys = [k_ for k_ in r0.files if k_ != "losses"]
    for key in keys:   # same values added in the same order on every rank
        np.testing.assert_array_equal(r0[key], r1[key], err_msg=key)
    # single GPU, whole batch, weights as rank 0 initialised them
    m = make_product_model(name, device="cuda:0", seed=2024, **SHAPE)
    batches, k = _batches(name, SHAPE)
    losses = []
    for b in batches:
        loss = m.calculate_loss(to_device_batch(b, "cuda:0"))
        loss.backward()
        losses.append(float(loss.item()))
    np.testing.assert_allclose(0.5 * (r0["losses"] + r1["losses"]), losses, rtol=1e-5)
    for key, v in m.state_dict().items():
        assert_weights_close(r0[key], v.cpu().numpy(), rtol=1e-5, atol=5e-7, err_msg=f"{name} {key}")


def _eval_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from hopwise_b200.distributed import reduce_metric_sums, shard_bounds
    from hopwise_b200.evaluator import topk_hits, topk_metric_sums

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dev = f"cuda:{rank}"
        m = make_product_model("DistMult", device=dev, **EVAL_SHAPE)
        users, hoff, hist, poff, pos = _eval_inputs()
        lo, hi = shard_bounds(len(users), rank, world)          # contiguous user block of this rank
        u = torch.from_numpy(users[lo:hi]).to(dev)
        ho = torch.from_numpy(hoff[lo:hi + 1] - hoff[lo]).to(dev)
        hi_items = torch.from_numpy(hist[hoff[lo]:hoff[hi]]).to(dev)
        po = torch.from_numpy(poff[lo:hi + 1] - poff[lo]).to(dev)
        pi = torch.from_numpy(pos[poff[lo]:poff[hi]]).to(dev)
        ids, _ = m.full_sort_topk(u, EVAL_K, ho, hi_items, return_scores=False)
        sums = topk_metric_sums(topk_hits(ids, po, pi))
        total, n = reduce_metric_sums(sums, hi - lo)            # NCCL all-reduce of the float64 sums + user counts
        np.savez(os.path.join(out_dir, f"eval{rank}.npz"), sums=total.cpu().numpy(), n=n, ids=ids.cpu().numpy())
    finally:
        dist.destroy_process_group()


EVAL_SHAPE = dict(U=1501, I=9001, E=9101, R=4, d=32)
EVAL_K = 20


def _eval_inputs():
    rng = np.random.default_rng(31)
    n = 1237                                                     # not divisible by the world size
    users = rng.integers(1, EVAL_SHAPE["U"], n)
    hlen = rng.integers(0, 40, n)
    hoff = np.concatenate([[0], np.cumsum(hlen)])
    hist = np.concatenate([np.sort(rng.choice(np.arange(1, EVAL_SHAPE["I"]), size=int(l), replace=False)) for l in hlen])
    plen = rng.integers(1, 6, n)
    poff = np.concatenate([[0], np.cumsum(plen)])
    pos = np.concatenate([np.sort(rng.choice(np.arange(1, EVAL_SHAPE["I"]), size=int(l), replace=False)) for l in plen])
    return users, hoff, hist, poff, pos


def test_two_rank_sharded_evaluation_matches_single_gpu(tmp_path):
    """users in contiguous blocks per rank, per-rank metric sums reduced over NCCL = the single-GPU
    evaluation of all users (exact global mean; the reference all-gathers rounded per-rank means)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from hopwise_b200.evaluator import metrics_from_sums, topk_hits, topk_metric_sums

    mp.spawn(_eval_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "eval0.npz"), np.load(tmp_path / "eval1.npz")
    np.testing.assert_array_equal(r0["sums"], r1["sums"])        # every rank holds the global sums
    users, hoff, hist, poff, pos = _eval_inputs()
    assert int(r0["n"]) == len(users)
    m = make_product_model("DistMult", device="cuda:0", **EVAL_SHAPE)
    dev = "cuda:0"
    ids, _ = m.full_sort_topk(torch.from_numpy(users).to(dev), EVAL_K, torch.from_numpy(hoff).to(dev),
                              torch.from_numpy(hist).to(dev), return_scores=False)
    np.testing.assert_array_equal(np.concatenate([r0["ids"], r1["ids"]]), ids.cpu().numpy())   # top-k ids: bit-exact
    want = topk_metric_sums(topk_hits(ids, torch.from_numpy(poff).to(dev), torch.from_numpy(pos).to(dev)))
    np.testing.assert_allclose(r0["sums"], want.cpu().numpy(), rtol=1e-12)
    a = metrics_from_sums(torch.from_numpy(r0["sums"]), len(users), [10, 20], decimals=None)
    b = metrics_from_sums(want, len(users), [10, 20], decimals=None)
    for key in b:
        assert abs(a[key] - b[key]) <= 1e-12
