"""Import the real hopwise from /root/reference (build container only).

Used by make_golden.py and by the few CPU tests that compare the oracle with the live
reference when it happens to be present.  Nothing that runs on the GPU box imports this.
"""

import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "hopwise"))


def import_reference():
    """Stub the three logging-only dependencies the container lacks, then import hopwise."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    if "colorama" not in sys.modules:
        m = types.ModuleType("colorama")
        m.init = lambda *a, **k: None
        sys.modules["colorama"] = m
    if "colorlog" not in sys.modules:
        m = types.ModuleType("colorlog")
        m.ColoredFormatter = type("ColoredFormatter", (), {"__init__": lambda self, *a, **k: None})
        sys.modules["colorlog"] = m
    if "texttable" not in sys.modules:
        m = types.ModuleType("texttable")
        m.Texttable = type("Texttable", (), {"__init__": lambda self, *a, **k: None})
        sys.modules["texttable"] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import hopwise  # noqa: F401

    return hopwise


class FakeDataset:
    """The attributes the four reference constructors and KGSampler read."""

    def __init__(self, n_users, n_items, n_entities, n_relations, heads=None, tails=None, ui_token="[UI-Relation]"):
        self._num = {
            "user_id": n_users,
            "item_id": n_items,
            "entity_id": n_entities,
            "relation_id": n_relations,
        }
        self.ui_relation = ui_token
        self.field2token_id = {"relation_id": {ui_token: n_relations - 1}}
        self.head_entity_field = "head_id"
        self.tail_entity_field = "tail_id"
        self.head_entities = heads
        self.tail_entities = tails
        self.entity_num = n_entities

    def num(self, field):
        return self._num[field]


REF_CONFIG = {
    "USER_ID_FIELD": "user_id",
    "ITEM_ID_FIELD": "item_id",
    "NEG_PREFIX": "neg_",
    "ENTITY_ID_FIELD": "entity_id",
    "RELATION_ID_FIELD": "relation_id",
    "HEAD_ENTITY_ID_FIELD": "head_id",
    "TAIL_ENTITY_ID_FIELD": "tail_id",
    "device": "cpu",
    "embedding_size": 16,
    "margin": 1.0,
}
