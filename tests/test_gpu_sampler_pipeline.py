"""The device samplers inside hopwise's OWN training loader (oracle/_ref): after
`hopwise_b200.sampler.install_device_samplers(train_data)` the KnowledgeBasedDataLoader of the reference yields,
batch for batch, the negatives its CPU samplers would have drawn -- same NumPy MT19937 stream (continued from
np.random.get_state()), same filtering, KG draws before rec draws (knowledge_dataloader.py:137-145) -- and the
sampler objects expose what the reference's loaders read (`phase`, `used_ids` as arrays of sets, set_phase)."""

import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref as oref  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not oref.ref_available(), reason="oracle/_ref is not built")]


def _pipeline(**extra):
    from hopwise.config import Config
    from hopwise.data import create_dataset, data_preparation
    from hopwise.utils import init_seed

    config = Config(model="TransE", dataset="ml-100k",
                    config_dict={"embedding_size": 16, "train_batch_size": 2048, "epochs": 1, "use_gpu": False, "gpu_id": "",
                                 "show_progress": False, "seed": 2024, **extra})
    init_seed(config["seed"], config["reproducibility"])
    dataset = create_dataset(config)
    train_data, valid_data, test_data = data_preparation(config, dataset)
    init_seed(config["seed"], config["reproducibility"])
    return config, train_data, valid_data, test_data


def test_device_samplers_reproduce_the_reference_loader(tmp_path):
    oref.import_ref()
    from hopwise.data.dataloader.knowledge_dataloader import KGDataLoaderState
    from hopwise_b200.sampler import KGSampler, RecSampler, install_device_samplers

    torch.zeros(1, device="cuda")   # CUDA context before hopwise's Config exports CUDA_VISIBLE_DEVICES = gpu_id
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        _, train_ref, _, _ = _pipeline()
        train_ref.set_mode(KGDataLoaderState.RSKG)
        want = [{k: v.clone() for k, v in b.interaction.items()} for b in train_ref]
        _, train_dev, valid_dev, _ = _pipeline()
        old_rec = train_dev.general_dataloader._sampler
        install_device_samplers(train_dev)
        rec, kgs = train_dev.general_dataloader._sampler, train_dev.kg_dataloader._sampler
        assert isinstance(rec, RecSampler) and isinstance(kgs, KGSampler)
        train_dev.set_mode(KGDataLoaderState.RSKG)
        got = [{k: v.clone() for k, v in b.interaction.items()} for b in train_dev]
    finally:
        os.chdir(cwd)
        if visible is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = visible
    assert len(got) == len(want) == 39
    for i, (a, b) in enumerate(zip(got, want)):
        assert set(a) == set(b)
        for k in b:
            assert not a[k].is_cuda, "hopwise's loader assembles the batch on the host"
            assert torch.equal(a[k], b[k]), f"batch {i} field {k}"
    # the attributes hopwise's loaders read from the sampler objects
    assert rec.phase == old_rec.phase == "train"
    ref_used = old_rec.used_ids
    mine = rec.used_ids
    assert len(mine) == len(ref_used)
    for u in (1, 7, 100, len(mine) - 1):
        assert mine[u] == set(int(x) for x in ref_used[u])
    # cumulative phases: the validation sampler forbids the train AND the validation items of a user
    ref_valid = valid_dev._sampler      # the reference's sampler bound to the validation phase (data/utils.py)
    assert ref_valid.phase == "valid"
    mine_valid = rec.set_phase("valid")
    assert mine_valid.phase == "valid" and mine_valid.stream is rec.stream
    for u in (1, 7, 100):
        assert mine_valid.used_ids[u] == set(int(x) for x in ref_valid.used_ids[u])
    with pytest.raises(ValueError):
        rec.set_phase("nope")
    kg_ref_used = train_ref.kg_dataloader._sampler.used_ids
    for h in (1, 50, 1000):
        assert kgs.used_ids[h] == set(int(x) for x in kg_ref_used[h])


@pytest.mark.parametrize("mode", ["popularity", "dynamic", "popularity+dynamic"])
def test_popularity_and_dynamic_sampling_inside_the_reference_loader(tmp_path, mode):
    """train_neg_sample_args with distribution: popularity (alpha 0.5) and / or dynamic: True (candidate_num 3):
    hopwise's own `_neg_sampling` (abstract_dataloader.py:166-183) draws the candidates from the sampler and keeps
    the one the model scores best.  With the device samplers installed -- and the fused model doing the scoring in
    both arms -- the loader yields the batches it yields with the reference's CPU samplers."""
    oref.import_ref()
    import hopwise_b200
    from hopwise.data.dataloader.knowledge_dataloader import KGDataLoaderState
    from hopwise_b200.sampler import install_device_samplers

    torch.zeros(1, device="cuda")
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    pop, dyn = "popularity" in mode, "dynamic" in mode
    args = {"train_neg_sample_args": {"distribution": "popularity" if pop else "uniform", "sample_num": 1,
                                      "alpha": 0.5 if pop else 1.0, "dynamic": dyn, "candidate_num": 3 if dyn else 0}}
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        batches = {}
        for arm in ("ref", "dev"):
            config, train, _, _ = _pipeline(**args)
            assert train.general_dataloader._sampler.distribution == ("popularity" if pop else "uniform")
            torch.manual_seed(7)
            model = hopwise_b200.TransE(config, train.dataset).to("cuda")
            model.device = torch.device("cuda")   # (the config says cpu: this pipeline keeps hopwise's loaders on the host)
            if arm == "dev":
                install_device_samplers(train)
                assert (train.kg_dataloader._sampler.pop is not None) == pop
            if dyn:
                train.get_model(model)
            train.set_mode(KGDataLoaderState.RSKG)
            got = []
            for i, b in enumerate(train):
                got.append({k: v.clone() for k, v in b.interaction.items()})
                if i == 5:
                    break
            batches[arm] = got
    finally:
        os.chdir(cwd)
        if visible is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = visible
    assert len(batches["ref"]) == len(batches["dev"]) == 6
    for i, (a, b) in enumerate(zip(batches["dev"], batches["ref"])):
        assert set(a) == set(b)
        for k in ("neg_tail_id", "head_id", "user_id", "item_id", "neg_item_id"):
            bad = (a[k] != b[k]).nonzero().flatten()
            assert bad.numel() == 0, f"batch {i} field {k}: {bad.numel()} of {a[k].numel()} differ, first at {bad[:5].tolist()}"
        for k in b:
            assert torch.equal(a[k], b[k]), f"batch {i} field {k}"
