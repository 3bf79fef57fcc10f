"""Shared helpers of the parity tests: build product models / oracle twins on given shapes."""

import numpy as np
import torch

from oracle.kge_torch import OracleKGE, Shapes

BATCH_KEYS = ("user_id", "item_id", "neg_item_id", "head_id", "relation_id", "tail_id", "neg_tail_id")

BASE_CONFIG = {
    "USER_ID_FIELD": "user_id",
    "ITEM_ID_FIELD": "item_id",
    "NEG_PREFIX": "neg_",
    "ENTITY_ID_FIELD": "entity_id",
    "RELATION_ID_FIELD": "relation_id",
    "HEAD_ENTITY_ID_FIELD": "head_id",
    "TAIL_ENTITY_ID_FIELD": "tail_id",
    "margin": 1.0,
    "learner": "adam",
    "learning_rate": 0.001,
    "weight_decay": 0.0,
}


class ShapeDataset:
    """The dataset attributes the model constructors read."""

    def __init__(self, U, I, E, R, ui_token_id=None):
        self._num = {"user_id": U, "item_id": I, "entity_id": E, "relation_id": R}
        self.ui_relation = "[UI-Relation]"
        self.field2token_id = {"relation_id": {"[UI-Relation]": R - 1 if ui_token_id is None else ui_token_id}}

    def num(self, field):
        return self._num[field]


def make_product_model(name, U, I, E, R, d, device="cuda", margin=1.0, seed=2024, lr=1e-3, **cfg):
    import hopwise_b200

    config = dict(BASE_CONFIG, embedding_size=d, margin=margin, device=torch.device(device), learning_rate=lr, **cfg)
    torch.manual_seed(seed)
    model = hopwise_b200.MODELS[name](config, ShapeDataset(U, I, E, R))
    return model.to(device)


def make_oracle_model(name, U, I, E, R, d, margin=1.0, seed=2024):
    torch.manual_seed(seed)
    return OracleKGE(name, Shapes(U, I, E, R, d, margin=margin))


def random_batch(rng, U, I, E, R, n_rec, n_kg, k_rec=1, k_kg=1):
    return {
        "user_id": rng.integers(1, U, n_rec),
        "item_id": rng.integers(1, I, n_rec),
        "neg_item_id": rng.integers(1, I, n_rec * k_rec),
        "head_id": rng.integers(1, E, n_kg),
        "relation_id": rng.integers(1, R - 1, n_kg),
        "tail_id": rng.integers(1, E, n_kg),
        "neg_tail_id": rng.integers(1, E, n_kg * k_kg),
    }


def to_device_batch(b, device="cuda"):
    return {k: torch.as_tensor(np.asarray(v), dtype=torch.long).to(device) for k, v in b.items()}


def tile_batch(b, k_rec, k_kg):
    """The reference's K-negative layout: positives repeated K times (abstract_dataloader.py:192-198)."""
    out = dict(b)
    for key in ("user_id", "item_id"):
        out[key] = np.tile(b[key], k_rec)
    for key in ("head_id", "relation_id", "tail_id"):
        out[key] = np.tile(b[key], k_kg)
    return out


def to_cpu_batch(b):
    return {k: torch.as_tensor(np.asarray(v), dtype=torch.long) for k, v in b.items()}


def assert_weights_close(got, want, rtol=1e-5, atol=5e-7, outlier_frac=5e-4, outlier_atol=1e-4, err_msg=""):
    """Post-Adam weights: every element within rtol/atol, except a bounded few.

    Adam's update lr * m / (sqrt(v) + eps) amplifies rounding noise of a near-zero gradient
    (|g| ~ eps = 1e-8, e.g. duplicate rows whose contributions cancel): d(update)/dg peaks at
    lr / (4 eps) = 2.5e4, so two correct fp32 implementations that sum in a different order can
    differ by up to ~1e-5..1e-4 on such an element.  Those elements are allowed, but only
    `outlier_frac` of a table (two at least) and never beyond `outlier_atol` (a tenth of one Adam step).
    """
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, err_msg
    diff = np.abs(got - want)
    bad = diff > atol + rtol * np.abs(want)
    assert np.isfinite(got).all(), err_msg
    # (at least two elements per table: the float atomics of the gradient scatter add in a different order every run, so
    # WHICH nearly cancelled element lands beyond the tolerance varies from run to run on small tables)
    allowed = max(2, int(outlier_frac * bad.size))
    assert bad.sum() <= allowed, f"{err_msg}: {bad.sum()} / {bad.size} elements beyond rtol={rtol}, atol={atol}"
    assert diff.max() <= outlier_atol, f"{err_msg}: max abs diff {diff.max()} > {outlier_atol}"
