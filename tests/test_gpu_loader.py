"""DeviceKGLoader on the GPU (MT19937 sampler kernels, device-side gathers) against the reference's batches:
every id vector of two ml-100k epochs bit-exact, final sampler state equal to numpy's."""

import zlib

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_device_kg_loader_gpu_matches_reference_epochs():
    from hopwise_b200.loader import DeviceKGLoader
    from hopwise_b200.sampler import KGSampler, MTStream, RecSampler

    g = load_golden("loader_ml100k.npz")
    stream = MTStream(state=("MT19937", g["mt_key"], int(g["mt_pos"])))
    n_users, n_items, n_ent = int(g["n_users"]), int(g["n_items"]), int(g["n_entities"])
    rec = RecSampler(g["used_user"].astype(np.int64), g["used_item"].astype(np.int64), n_users, n_items, stream=stream)
    kg = KGSampler(heads=g["sampler_heads"].astype(np.int64), tails=g["sampler_tails"].astype(np.int64), entity_num=n_ent,
                   stream=stream)
    loader = DeviceKGLoader(g["inter_user"], g["inter_item"], g["kg_head"], g["kg_rel"], g["kg_tail"], rec, kg,
                            batch_size=int(g["batch"]), seed=int(g["seed"]), device="cuda")
    for ep in range(2):
        steps = 0
        for i, b in enumerate(loader):
            for c, key in enumerate(DeviceKGLoader.KEYS):
                assert b[key].is_cuda and b[key].dtype == torch.int64
                v = b[key].cpu().numpy()
                assert zlib.crc32(v.tobytes()) == int(g["crc"][ep, i, c]), (ep, i, key)
            steps += 1
        assert steps == 39
    state = stream.get_state()
    np.testing.assert_array_equal(state[1], g["mt_key_end"])
    assert state[2] == int(g["mt_pos_end"])


def test_device_kg_loader_feeds_the_fused_step():
    """A short training run driven by the device loader: finite, decreasing loss."""
    from hopwise_b200.loader import DeviceKGLoader
    from hopwise_b200.sampler import KGSampler, MTStream, RecSampler
    from kge_helpers import make_product_model

    g = load_golden("loader_ml100k.npz")
    stream = MTStream(seed=2024)
    n_users, n_items, n_ent, n_rel = (int(g[k]) for k in ("n_users", "n_items", "n_entities", "n_relations"))
    rec = RecSampler(g["used_user"].astype(np.int64), g["used_item"].astype(np.int64), n_users, n_items, stream=stream)
    kg = KGSampler(heads=g["sampler_heads"].astype(np.int64), tails=g["sampler_tails"].astype(np.int64), entity_num=n_ent,
                   stream=stream)
    loader = DeviceKGLoader(g["inter_user"], g["inter_item"], g["kg_head"], g["kg_rel"], g["kg_tail"], rec, kg,
                            batch_size=2048, seed=2024)
    m = make_product_model("TransE", n_users, n_items, n_ent, n_rel, 64, lr=1e-2)   # relation_num counts [UI-Relation]
    epoch_loss = []
    for _ in range(3):
        tot = 0.0
        for b in loader:
            loss = m.calculate_loss(b)
            tot += float(loss.item())
            loss.backward()
        epoch_loss.append(tot)
    assert np.isfinite(epoch_loss).all() and epoch_loss[-1] < epoch_loss[0]
