"""Import the real hopwise (oracle/_ref, built from /root/reference by oracle/build_ref.py).

Used by make_golden.py and by the tests that compare the oracle or the product with the live
reference.  Test infrastructure only.
"""

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def reference_available() -> bool:
    from oracle import ref

    return ref.ref_available()


def import_reference():
    """hopwise from oracle/_ref: the unmodified copy of /root/reference/hopwise that oracle/build_ref.py makes (plus
    stubs for the three logging-only dependencies the image lacks).  One copy per process, whoever asks first."""
    from oracle import ref

    return ref.import_ref()


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
